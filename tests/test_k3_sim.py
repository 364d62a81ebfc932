"""k2a_v3's warp body executed on the CPU (tests/cpp/k3_sim.cu: the same __host__ __device__ template the
kernel instantiates, 32 host threads per warp) against a plain restatement of vfo::process's front end
(vfo.cpp:237-251, halfbanddecimator.cpp:43-72, dsp.cpp:163-173): table wrap, start-up transient, stream
sample 0, callback heads, ragged stream groups, several spans per callback, state carried over three calls."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, ".."))
EXE = os.path.join(HERE, "cpp", "k3_sim")


@pytest.fixture(scope="module")
def sim(built):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = os.path.join(HERE, "cpp", "k3_sim.cu")
    hdr = os.path.join(ROOT, "sdrreceiver_b200", "csrc", "kernels_v3.cuh")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        if not os.path.exists(nvcc):
            pytest.skip("nvcc not available")
        r = subprocess.run([nvcc, "-std=c++20", "-O1", "-w", "-gencode", "arch=compute_100a,code=sm_100a", "-o", EXE, src,
                            os.path.join(ROOT, "sdrreceiver_b200", "csrc", "plan_host.o"), "-Xcompiler", "-pthread"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    return EXE


# fs, callback size, streams, streams per warp, spans per callback, sub VFOs
@pytest.mark.parametrize("args", [
    (30720, 7680, 3, 2, 4, 6),       # two stream groups, the second one half empty
    (30720, 7680, 5, 3, 1, 6),       # odd streams per warp (a pair with one member), one span
    (30720, 7680, 4, 2, 7, 16),      # 16 sub VFOs, ragged last span
    (30720, 7680, 3, 8, 3, 4),       # eight streams per warp
    (30720, 7680, 2, 1, 5, 16),      # one stream per warp
    (30720, 7680, 3, 2, 4, 6, 1),    # the warp-specialised producer / consumer pair, two ring buffers
    (30720, 7680, 4, 2, 7, 16, 1),
    (30720, 7680, 5, 3, 1, 6, 1),
])
def test_k3_unit_matches_the_plain_cascade(sim, args):
    r = subprocess.run([sim] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:]
