"""In-tree build of the native libraries.

    librayito_b200.so   CUDA render core + C ABI      (rayito_b200/csrc, nvcc, sm_100a only)
    librayito_host.so   C++ mirror of the Rayito API  (rayito_b200/host, g++)
    fixtures/librayito_fixtures.so   recipe scenes of the tests and of bench.py (not product)

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container;
the resulting .so files are git-ignored but travel to the GPU box with the snapshot.
The system compiler is named explicitly: the image exports CXX=/opt/gcc/bin/g++, a -B
wrapper whose shared objects crash when dlopen()ed from Python.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "rayito_b200", "csrc")
HOST = os.path.join(ROOT, "rayito_b200", "host")
INCLUDE = os.path.join(ROOT, "include")
CORE_LIB = os.path.join(CSRC, "librayito_b200.so")
HOST_LIB = os.path.join(HOST, "librayito_host.so")
FIXTURES = os.path.join(ROOT, "fixtures")
FIXTURES_LIB = os.path.join(FIXTURES, "librayito_fixtures.so")
ASSETS = os.path.join(ROOT, "assets", "_models")
REFERENCE = "/root/reference"

HOST_CXX = "/usr/bin/g++"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# -fmad=false is a correctness flag, not a tuning flag: hit decisions must be
# bit-identical to the reference's unfused x86 arithmetic.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-ccbin", HOST_CXX,
    "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(directory, exts):
    out = []
    for base, _dirs, files in os.walk(directory):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return out


def _run(cmd, cwd=None):
    proc = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return proc.stdout


def build_core(force=False, verbose=False):
    srcs = _sources(CSRC, (".cu", ".cuh", ".h")) + _sources(INCLUDE, (".h",))
    if force or _newer(CORE_LIB, srcs):
        # RT_NVCC_EXTRA: extra -D switches for tuning runs (e.g. "-DRT_ADVANCE_STEPS=2")
        extra = os.environ.get("RT_NVCC_EXTRA", "").split()
        cmd = [NVCC] + NVCC_FLAGS + extra + ["-I" + INCLUDE, os.path.join(CSRC, "rt_core.cu"), "-o", CORE_LIB]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        out = _run(cmd)
        if verbose:
            print(out)
    return CORE_LIB


def build_host(force=False):
    build_core()
    srcs = _sources(HOST, (".cpp", ".hpp", ".h")) + _sources(INCLUDE, (".h",))
    if force or _newer(HOST_LIB, srcs + [CORE_LIB]):
        # -ffp-contract=off and no -march: host float arithmetic feeds BVH builds and
        # transform keys and must equal the reference's.
        cmd = [HOST_CXX, "-O2", "-std=c++11", "-fPIC", "-shared", "-ffp-contract=off", "-Wall",
               "-I" + HOST, "-I" + INCLUDE,
               os.path.join(HOST, "host_impl.cpp"), "-o", HOST_LIB,
               "-L" + CSRC, "-lrayito_b200", "-pthread", "-Wl,-rpath,$ORIGIN/../csrc", "-Wl,-Bsymbolic"]
        _run(cmd)
    return HOST_LIB


def build_fixtures(force=False):
    """fixtures/librayito_fixtures.so: the recipe scenes of the tests and of bench.py, built with
    the host library's public API (test infrastructure, not product)."""
    build_host()
    srcs = _sources(FIXTURES, (".cpp", ".h")) + _sources(HOST, (".hpp", ".h")) + _sources(INCLUDE, (".h",))
    if force or _newer(FIXTURES_LIB, srcs + [HOST_LIB]):
        cmd = [HOST_CXX, "-O2", "-std=c++11", "-fPIC", "-shared", "-ffp-contract=off", "-Wall",
               "-I" + FIXTURES, "-I" + HOST, "-I" + INCLUDE,
               os.path.join(FIXTURES, "fixtures.cpp"), "-o", FIXTURES_LIB,
               "-L" + HOST, "-lrayito_host", "-L" + CSRC, "-lrayito_b200", "-pthread",
               "-Wl,-rpath,$ORIGIN/../rayito_b200/host", "-Wl,-rpath,$ORIGIN/../rayito_b200/csrc", "-Wl,-Bsymbolic"]
        _run(cmd)
    return FIXTURES_LIB


def stage_assets():
    """Copy the reference's OBJ fixtures into assets/_models (git-ignored) when the
    reference tree is present; the GPU box has no /root/reference and uses the copy."""
    src_dir = os.path.join(REFERENCE, "models")
    if not os.path.isdir(src_dir):
        return
    os.makedirs(ASSETS, exist_ok=True)
    for name in ("bumpy.obj",):
        src, dst = os.path.join(src_dir, name), os.path.join(ASSETS, name)
        if os.path.exists(src) and (not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src)):
            with open(src, "rb") as f, open(dst, "wb") as g:
                g.write(f.read())


def build_oracle():
    """Compile the checker: the unmodified reference (oracle/_ref, only when
    /root/reference exists) and the C restatement (oracle/libport)."""
    oracle = os.path.join(ROOT, "oracle")
    if os.path.isdir(REFERENCE):
        _run(["make", "-C", oracle, "ref"])
    _run(["make", "-C", oracle, "port"])


def build_all(force=False):
    stage_assets()
    build_core(force)
    build_host(force)
    build_fixtures(force)
    build_oracle()


def model_path(name="bumpy.obj"):
    for candidate in (os.path.join(ASSETS, name), os.path.join(REFERENCE, "models", name)):
        if os.path.exists(candidate):
            return candidate
    raise FileNotFoundError(
        name + " not found: run __graft_entry__.build() where /root/reference exists to stage it under assets/_models")


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", CORE_LIB, HOST_LIB)
