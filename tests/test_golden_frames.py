"""Golden frame sequences made by the reference's own shader sources (tests/make_golden_frames.py, committed under
tests/golden/frames_*.npz): the oracle must reproduce them on the CPU, the CUDA path (through the C ABI) on the GPU.
Bit-exact on every reservoir field; the colour output bit-exact too (gamma 1).  These run where neither
/root/reference nor oracle/_ref exist."""
import ast
import glob
import os

import numpy as np
import pytest

import make_golden_frames as mg
import parity_harness as ph

FILES = sorted(glob.glob(os.path.join(mg.GOLDEN, "frames_*.npz")))


def _load(path):
    z = np.load(path)
    spec = ast.literal_eval(str(z["spec"]))
    case = mg.build_case(spec)
    assert mg.gbuffer_checksum(case) == int(z["gbuffer_crc32"]), "fixture inputs changed: regenerate with tests/make_golden_frames.py"
    return z, spec, case


def _check(frames, z, label):
    for f, fr in enumerate(frames):
        n = ph.compare_reservoirs(fr["initial"], z[f"initial_{f}"], f"{label} frame {f} after restirOmni")
        assert n == 0, f"{label} frame {f}: {n} reservoirs differ from the golden after restirOmni"
        n = ph.compare_reservoirs(fr["reservoirs"], z[f"final_{f}"], f"{label} frame {f} final")
        assert n == 0, f"{label} frame {f}: {n} final reservoirs differ from the golden"
        eq = ph.bits_equal(fr["rgba"][..., :3], z[f"rgba_{f}"][..., :3])
        assert eq.all(), f"{label} frame {f}: {(~eq).sum()} colour values differ from the golden"


def test_golden_fixtures_present():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p)[:-4] for p in FILES])
def test_oracle_reproduces_reference_made_frames(path):
    z, spec, case = _load(path)
    _check(ph.run_oracle(case), z, "oracle")


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p)[:-4] for p in FILES])
def test_cuda_reproduces_reference_made_frames(path):
    z, spec, case = _load(path)
    _check(ph.run_cuda(case), z, "cuda")
