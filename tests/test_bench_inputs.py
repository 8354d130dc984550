"""bench.py's CPU arm (`--impl reference`) and its inputs (oracle/ref_inputs.py): runs without the product package and
without librestir_b200.so, on the reference's own blobs; host logic only, no GPU."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import parity_harness as ph

ROOT = ph.ROOT


def _ref_inputs():
    spec = importlib.util.spec_from_file_location("ref_inputs", os.path.join(ROOT, "oracle", "ref_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", ["cornellBox", "sponza", "office"])
def test_reference_scene_blobs_and_material_table(name):
    """The reference arm's scene is the reference's own output, and its material table (stated independently in
    oracle/ref_inputs.py) equals the product's (fixtures.material_table) for every material of the three scenes —
    metallic-roughness and specular-glossiness (office) alike."""
    ri = _ref_inputs()
    if not ri.available(name):
        pytest.skip("scenes/_baked missing")
    ref = ri.ReferenceScene(name)
    mine = ph.fixtures.load_baked(name, rebuild=True)
    assert np.array_equal(ref.material_table(), mine.material_table())
    assert np.array_equal(ref.nodes, mine.nodes) and np.array_equal(ref.alias_blob, mine.alias_blob)   # host builders == reference's
    assert ref.light_counts() == mine.light_counts()


@pytest.mark.parametrize("config,gpus", [("cornell_tiny_biased4", 1), ("cornell_tiny_unbiased3", 1), ("cornell_tiny_unbiased3", 2)])
def test_reference_arm_runs_without_the_product(config, gpus):
    ri = _ref_inputs()
    if not ri.available("cornellBox"):
        pytest.skip("scenes/_baked missing")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", config, "--gpus", str(gpus),
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, check=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "Mrays/s" and line["higher_is_better"] is True
    assert line["product_modules_loaded"] == []
    assert all(so.startswith("oracle/") for so in line["native_so_loaded"]), line["native_so_loaded"]
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0
    if gpus == 1:
        w, h = (160, 90) if "biased4" in config else (160, 96)
        assert line["ms_per_frame"] == line["ms_per_step"] and f"{w}x{h}" in line["config"]["workload"]
        if "unbiased" in config:
            assert w * h * 2 <= line["rays_per_step"] <= w * h * 5
        else:
            assert line["rays_per_step"] == w * h
    else:
        assert "3 windows" in line["config"]["workload"] and line["ms_per_frame"] is None
