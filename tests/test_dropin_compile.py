"""The host headers are a drop-in for tutorial code: a Stage 6 style program compiles against
them unchanged with -DRAYITO_B200_STAGE=6 and gets Stage 6 rules (no GPU needed up to
prepare()/flatten)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stage6_application_builds_against_dropin_headers(tmp_path, obj_path, capi):
    host = os.path.join(ROOT, "rayito_b200", "host")
    core = os.path.join(ROOT, "rayito_b200", "csrc")
    exe = str(tmp_path / "stage6_app")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++11", "-DRAYITO_B200_STAGE=6", "-I" + host, "-I" + os.path.join(ROOT, "fixtures"),
                    "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "stage6_app.cpp"),
                    "-L" + host, "-lrayito_host", "-L" + core, "-lrayito_b200",
                    "-Wl,-rpath," + host, "-Wl,-rpath," + core, "-o", exe], check=True, timeout=300)
    out = subprocess.run([exe, obj_path], check=True, capture_output=True, text=True, timeout=120).stdout
    assert "semantics=6 shapes=9 lights=2 top_nodes=15 shutter=0,0" in out, out
