"""Host-side pieces of bench.py that the GPU box only exercises on one branch: the clock sampler's `nvidia-smi` fallback
(the box has the NVML module, so its runs take the in-process path)."""
import datetime
import os
import sys
import tempfile
import time

from conftest import ROOT

sys.path.insert(0, ROOT)


def test_clock_sampler_fallback_uses_only_samples_after_the_mark():
    import bench
    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.nv = None
    s.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)

    class Done:
        def terminate(self): pass
        def wait(self): pass
        def poll(self): return 0
    s.p = Done()
    now = time.time()

    def line(t, mhz, power_cap):
        ts = datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        return "%s, %d, 1965, 500.0, 0x0, Not Active, Not Active, Not Active, %s\n" % (ts, mhz, "Active" if power_cap else "Not Active")
    s.f.write(line(now - 5.0, 1200, True))      # warm-up sample: ignored
    s.f.write(line(now + 0.2, 1950, False))
    s.f.write(line(now + 0.7, 1965, False))
    s.f.write("garbage line\n")
    s.t_mark = now
    r = s.stop()
    assert r["samples"] == 2 and r["sm_mhz"] == 1957.5 and r["sm_max_mhz"] == 1965 and r["reasons"] == []
    assert not os.path.exists(s.f.name)


def test_clock_sampler_without_any_tool():
    import bench
    c = bench.ClockSampler(0)
    c.wait_ready(0.1)
    c.mark()
    r = c.stop()
    assert "sm_mhz" in r and "reasons" in r
