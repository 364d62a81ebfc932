"""bench.py's reference arm runs on host cores only, so its JSON line can be checked here without a GPU:
the keys the driver reads (metric, value, unit, steps, e2e, cpu_baseline, impl) and that the arm never
touches a GPU (gpu_launches 0). The B200 arm's line is produced on the GPU box; its schema is checked by
tests/test_parity_gpu.py::test_bench_line_on_the_gpu."""
import json
import os
import subprocess
import sys

import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(out):
    lines = [x for x in out.splitlines() if x.startswith("{")]
    assert len(lines) == 1, out[-2000:]
    return json.loads(lines[0])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference once)")
def test_reference_arm_json_line():
    # SDRB_LIB points nowhere: the arm must not load the product library at all (it would fail loudly if it tried)
    env = dict(os.environ, SDRB_LIB="/nonexistent/libsdrb200.so")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1", "--blocks", "1"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _line(r.stdout)
    assert d["impl"] == "reference" and d["unit"] == "MS/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("aggregate input MS/s") and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["config"]["plan"] == "25E" and "workload" in d["config"]


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference once)")
def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29599")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                        "--blocks", "1"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [x for x in r.stdout.splitlines() if x.startswith("{")]
