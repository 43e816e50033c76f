"""Runs the C++17 QA executables of the host boundary (tests/cpp): the gr::block / gr::buffer /
edge::set_custom_buffer / flowgraph / scheduler_mt compatible harness on CPU, and -- on the GPU
box -- the B200 blocks + device-resident buffers inside real flowgraphs, mirroring the
reference's own scheduler tests."""
import os
import subprocess

import pytest

from conftest import ROOT

CPP = os.path.join(ROOT, "tests", "cpp")


def _build(target):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    subprocess.check_call(["make", "-C", CPP, "-s", f"build/{target}"])
    return os.path.join(CPP, "build", target)


def _run(exe, timeout):
    p = subprocess.run([exe], capture_output=True, text=True, timeout=timeout)
    print(p.stdout[-4000:])
    print(p.stderr[-2000:])
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-1000:]
    assert " 0 failed" in p.stdout
    return p.stdout


def test_host_boundary_cpu():
    out = _run(_build("qa_host_cpu"), 300)
    for name in ("SchedulerMTTest.TwoSinks", "SchedulerMTTest.BlockFanout",
                 "SchedulerBlockGrouping.BasicBlockGrouping", "SchedulerMTTags.PropagationPolicies",
                 "SchedulerMTTags.DecimationScalesOffsets", "SchedulerMTTest.BlockExceptionSurfacesInWait"):
        assert f"[  OK  ] {name}" in out


def test_end_of_stream_drain_is_race_free():
    """The end-of-stream decision for a 0/0 work() call raced with a draining reader (1 failing run in 8 before the
    fix in scheduler_mt.hpp): the interpolating-block QA, many times over."""
    exe = _build("qa_host_cpu")
    for _ in range(25):
        p = subprocess.run([exe, "OutputMultipleDrain"], capture_output=True, text=True, timeout=120)
        assert p.returncode == 0 and " 0 failed" in p.stdout, p.stdout[-2000:] + p.stderr[-500:]


@pytest.mark.gpu
def test_flowgraphs_on_gpu():
    out = _run(_build("qa_cuda_flowgraph"), 600)
    for name in ("SchedulerMTTest.CudaCopyBasic", "SchedulerMTTest.CudaCopyMultiThreaded",
                 "SchedulerMTTest.CudaCopyPinnedBuffers",
                 "Config1.FirCcf64", "Config1.FirFff64TensorCore", "Config2.FftMag", "Config3.FirMulFftChain", "Config4.PfbChannelizer64",
                 "Config4.PfbChannelizer64TensorCoreDft",
                 "Fusion.AdjacentBlocksCollapse", "TwoInput.MultiplyAndAdd",
                 "SchedulerMTTags.TagsAcrossDeviceBuffers"):
        assert f"[  OK  ] {name}" in out
