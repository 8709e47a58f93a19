"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`: the restated CPU baseline, rank 0
only under torchrun) prints exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = _run({"OMP_NUM_THREADS": "1"})          # what torchrun exports; the baseline must still use every core
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is False and d["unit"] == "s/molecule"
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["config"]["workload"] == "tiny" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == len(os.sched_getaffinity(0))
    assert "NOT votca/xtp" in d["config"]["note"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}) == []


def test_main_arm_assembles_the_json_line_with_a_fake_device(monkeypatch, capsys):
    """bench.main() end to end on a CPU box: the device (torch.cuda, the job that drives libxtpb200) is faked, everything
    else -- warm-up/timed-region bookkeeping, profiler and allocation statistics (host code of the real library),
    roofline/traffic assembly, e2e leg with its pinned-buffer lifetime, the sampled CPU baseline -- is the real code.
    Guards the keys the driver reads and the order pin -> e2e -> unpin."""
    import importlib.util
    import time
    import types

    import numpy as np
    import torch

    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from xtp_b200 import synth

    class FakeEvent:
        def __init__(self, enable_timing=False):
            self.t = 0.0

        def record(self):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    calls = []

    class FakeJob:
        def __init__(self, workload, device, rank=0, world=1, comm=None, e2e=True, seed=None, sigma="ppm", evgw=1):
            self.sz = synth.WORKLOADS[workload]
            self.pk = self.sz.n_basis * (self.sz.n_basis + 1) // 2
            self.last = {}
            self.pinned = False

        def run(self, resident=True):
            assert resident or self.pinned, "the e2e leg needs the pinned host copy"
            assert not (resident and self.pinned), "the resident leg must not see the pinned buffer"
            calls.append("resident" if resident else "e2e")
            sz = self.sz
            self.last = {"qp": np.linspace(-1, 1, sz.qptotal), "singlets": np.array([0.3, 0.4]), "vectors": None,
                         "davidson_info": "Success", "davidson_iterations": 7, "qp_unconverged": 0,
                         "stage_seconds": {"fill3c": 0.01, "bse_davidson": 0.005, "total": 0.02},
                         "grid_scan": {"compressed": True, "bins": 30, "direct_evaluations": 10.0,
                                       "equivalent_evaluations": 100.0}}
            time.sleep(0.002)
            return self.last

        def solve_bse_again(self):
            calls.append("bse-factorised")
            return {"singlets": np.array([0.3, 0.4]), "iterations": 9}

        def pin_host_copy(self):
            calls.append("pin")
            self.pinned = True

        def unpin_host_copy(self):
            calls.append("unpin")
            self.pinned = False

        def h2d_bytes(self, resident):
            return 1000 if resident else 5000

        def d2h_bytes(self):
            return 300

        def close(self):
            pass

    class FakeSampler:
        def __init__(self, dev):
            pass

        def start(self):
            pass

        def stop(self):
            return {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "power_w_max": 700.0, "samples": 3}

    from xtp_b200 import api
    # profile_get synchronises the device; with no device there were no launches to report either
    monkeypatch.setattr(api, "profile_summary", lambda: {})
    monkeypatch.setattr(bench, "GwbseJob", FakeJob)
    monkeypatch.setattr(bench, "ClockSampler", FakeSampler)
    monkeypatch.setattr(bench.sys, "argv", ["bench.py", "--workload", "tiny", "--steps", "2", "--warmup", "3"])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "XTPB_ALLOC_CACHE", "XTPB_BENCH_CACHE_DEFAULTED"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("NCCL_DEBUG", "WARN")
    bench.main()
    monkeypatch.delenv("XTPB_ALLOC_CACHE", raising=False)            # main() turns it on for single-GPU runs
    monkeypatch.delenv("XTPB_BENCH_CACHE_DEFAULTED", raising=False)
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks",
              "host_alloc"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["dtype"] == "f64" and d["value"] > 0
    assert d["config"]["workload"] == "tiny" and "model" not in d["config"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert d["roofline"]["traffic"] == d["roofline"]["traffic_detail"]["dram_bytes_per_launch"] > 1e9
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["e2e"]["h2d_bytes_per_step"] == 5000 and d["e2e"]["d2h_bytes_per_step"] == 300 and d["e2e"]["value"] > 0
    assert d["host_alloc"]["block_cache"] is True
    # 3 warm-up + 2 timed resident steps, then pin -> (1 warm-up + 2 timed) e2e steps -> unpin
    assert calls == ["resident"] * 5 + ["bse-factorised", "pin"] + ["e2e"] * 3 + ["unpin"]
    assert d["bse_modes"]["factorised_iterations"] == 9 and d["bse_modes"]["default"] == "dense"
