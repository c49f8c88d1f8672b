"""Config C1: the Stage 1 program.  Golden vector: Rayito_Stage1/out_ref.ppm (kept as
digest + structure in tests/golden/stage1_out_ref.json).  The oracle (the reference's
own main.cpp compiled where it lies) must reproduce it byte for byte, and so must the
GPU path through the C ABI -- 0 differing bytes, not merely <= 1/255."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "stage1_out_ref.json")))
STAGE1_BIN = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "stage1")


def _expected_payload():
    img = np.zeros((GOLD["height"], GOLD["width"], 3), np.uint8)
    img[GOLD["first_lit_row"]:GOLD["last_lit_row"] + 1] = GOLD["lit_colour"]
    assert hashlib.md5(img.tobytes()).hexdigest() == GOLD["payload_md5"]
    return img


def test_oracle_stage1_reproduces_golden(tmp_path):
    if not os.path.exists(STAGE1_BIN):
        pytest.skip("oracle/_ref/stage1 not built")
    subprocess.run([STAGE1_BIN], cwd=str(tmp_path), check=True, timeout=60)
    data = open(tmp_path / "out.ppm", "rb").read()
    assert len(data) == GOLD["bytes"]
    assert hashlib.md5(data).hexdigest() == GOLD["md5"]


@pytest.mark.gpu
def test_gpu_stage1_matches_golden(capi):
    img = capi.stage1_render(GOLD["width"], GOLD["height"])
    want = _expected_payload()
    assert int((img != want).sum()) == 0
    ppm = GOLD["header"].encode() + img.tobytes()
    assert hashlib.md5(ppm).hexdigest() == GOLD["md5"]


@pytest.mark.gpu
def test_gpu_stage1_other_sizes_match_oracle_rule(capi):
    # pixel-corner rays: the horizon row is where yu crosses 0.5, for any size
    for (w, h) in ((64, 64), (33, 17), (2, 2)):
        img = capi.stage1_render(w, h)
        yu = 1.0 - np.arange(h, dtype=np.float32) / np.float32(h - 1)
        lit = yu < 0.5
        assert (img[lit] == GOLD["lit_colour"]).all() and not img[~lit].any()
