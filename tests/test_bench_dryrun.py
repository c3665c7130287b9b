"""bench.py end to end WITHOUT a GPU: torch.cuda, NVML and the device-side tracer are replaced by stand-ins so that the
whole control flow of the N = 1 run (device-timed leg, roofline leg, end-to-end leg, reference call pattern, CPU
baseline, JSON line) executes on the CPU box. What is checked is the shape of the line the driver parses — every key
of the benchmark contract — and the bookkeeping (ray counts, launch counts, roofline arithmetic); the numbers
themselves are meaningless here. The real run happens on a B200 (`python bench.py`); this test only makes sure a change
to bench.py cannot break the contract unnoticed."""
import contextlib
import importlib.util
import io
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRIMARY, SHADOW = 2073600, 525594  # thai2 1080p pinned frame


class FakeTensor:
    def __init__(self, arr):
        self.a = np.asarray(arr)

    def pin_memory(self):
        return self

    def data_ptr(self):
        return self.a.ctypes.data

    def numpy(self):
        return self.a

    def zero_(self):
        return self

    def view(self, dt):
        return self

    def __float__(self):
        return float(self.a.reshape(-1)[0])

    def __getitem__(self, i):
        return self.a[i]


class FakeEvent:
    clock = [0.0]

    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        FakeEvent.clock[0] += 0.05
        self.t = FakeEvent.clock[0]

    def elapsed_time(self, other):
        return other.t - self.t


def fake_torch():
    cuda = types.SimpleNamespace(
        is_available=lambda: True, set_device=lambda d: None, synchronize=lambda d=None: None,
        Stream=lambda device=None: types.SimpleNamespace(cuda_stream=0), stream=lambda s: contextlib.nullcontext(), Event=FakeEvent,
        get_device_properties=lambda d: types.SimpleNamespace(multi_processor_count=148))
    t = types.ModuleType("torch")
    t.cuda, t.uint8, t.int32, t.int64, t.float64 = cuda, np.uint8, np.int32, np.int64, np.float64
    t.device = lambda kind, idx=0: (kind, idx)
    t.empty = lambda n, dtype=None, device=None: FakeTensor(np.zeros(min(int(n), 3840 * 2160), np.uint32))
    t.tensor = lambda v, dtype=None, device=None: FakeTensor(np.array(v))
    dist = types.ModuleType("torch.distributed")
    dist.init_process_group = lambda *a, **k: None
    dist.destroy_process_group = lambda: None
    dist.barrier = lambda: None
    dist.all_reduce = lambda tensor, op=None: None  # one rank's view: the other rank's numbers are not added
    dist.ReduceOp = types.SimpleNamespace(MAX="max")
    t.distributed = dist
    return t, dist


def fake_pynvml():
    m = types.ModuleType("pynvml")
    m.nvmlInit = lambda: None
    m.nvmlDeviceGetHandleByIndex = lambda i: i
    m.NVML_CLOCK_SM = 0
    m.nvmlDeviceGetMaxClockInfo = lambda h, c: 1965
    m.nvmlDeviceGetClockInfo = lambda h, c: 1965
    m.nvmlDeviceGetCurrentClocksThrottleReasons = lambda h: 0
    for k, name in enumerate(["SwPowerCap", "HwSlowdown", "SwThermalSlowdown", "HwThermalSlowdown", "HwPowerBrakeSlowdown", "SyncBoost",
                              "ApplicationsClocksSetting"]):
        setattr(m, "nvmlClocksThrottleReason" + name, 1 << k)
    return m


class FakeTracer:
    """Counts what bench.py asks of the tracer; a frame 'renders' PRIMARY + SHADOW rays per full-frame sample."""

    def __init__(self, w, h, recursions=0):
        self.w, self.h, self.rows_per_call, self.cur, self.recursions = w, h, 50, 0, recursions
        self.primary = self.shadow = self.bounce = self.kernels = 0
        self.tuning, self.calls = {}, []
        self.camera = types.SimpleNamespace(set_state=lambda *a: self.calls.append("camera"), move_rel=lambda *a: self.calls.append("move"))
        self.film = types.SimpleNamespace(clear=lambda: self.calls.append("clear"), pixel_datas=lambda: np.zeros((w * h, 7), np.float32))

    def _trace(self, rows, spp):
        time.sleep(0.002)  # the clock sampler takes a sample every 2 ms: the timed region must see a few
        frac = rows * spp / self.h
        self.primary += int(PRIMARY * frac)
        self.shadow += int(SHADOW * frac)
        if self.recursions:
            self.bounce += int(SHADOW * frac)
        self.kernels += 1

    def set_tuning(self, k, v):
        self.tuning[k] = v

    def set_stream(self, s):
        pass

    def configure(self, **kw):
        self.calls.append(("configure", kw.get("jitter_mode")))

    def trace_rows(self, first, n, spp=1, want_shadow=True):
        self._trace(n, spp)
        return n * self.w * spp, None

    def trace_frame_additive(self):
        self._trace(self.rows_per_call, 1)
        return self.rows_per_call * self.w

    def set_rows_per_call(self, r):
        self.rows_per_call = r

    def ray_totals(self):
        return {"primary": self.primary, "shadow": self.shadow, "bounce": self.bounce}

    def kernels_launched(self):
        return self.kernels

    def launch_stats(self):
        return {"trace_kernel_ms": 0.15, "kernels_launched": 1}

    def sync_timeouts(self):
        return 0

    def wait_pixels(self, keep=0):
        self.calls.append("wait%d" % keep)

    def get_tonemapped_pixels_async(self, ptr):
        self.calls.append("async")

    def get_tonemapped_pixels_into(self, ptr):
        self.calls.append("readback")

    def get_tonemapped_pixels_delta_into(self, ptr):
        self.calls.append("delta")

    def get_tonemapped_pixels(self, out=None):
        if out is None:
            out = np.zeros(self.w * self.h, np.uint32)
        out[:] = 0
        return out

    def get_primary_ids(self):
        return np.zeros(self.w * self.h, np.uint32)

    def set_host_frame(self, p):
        pass

    def close(self):
        self.calls.append("close")


class FakeGather:
    """Stands in for multi_gpu.FrameGather (which needs CUDA IPC): records the order of bench.py's calls."""

    last = None
    all = []

    def __init__(self, tracer, rank, world, dev, stream, mode="peer", fused_signal=False, fence="kernel"):
        self.kernels, self.calls, self.mode, self.fused = 0, [], mode, fused_signal
        FakeGather.last = self
        FakeGather.all.append(self)

    def begin_frame(self):
        self.calls.append("begin")

    def device_gather(self, release=False):
        self.kernels += 1
        self.calls.append("gather_release" if release else "gather")

    def read_frame_async(self, host):
        self.calls.append("read_async")

    def wait_frame(self, keep=0):
        self.calls.append("wait%d" % keep)

    def read_frame_into(self, host):
        self.calls.append("read_sync")

    def rearm(self):
        self.calls.append("rearm")

    def close(self):
        self.calls.append("close")


class FakeHostGather:
    """Stands in for multi_gpu.HostFrameGather (shared memory + cudaHostRegister)."""

    all = []

    def __init__(self, tracer, rank, world, dev, stream, name):
        self.calls, self.kernels, self.frame_no, self.w, self.h, self.name = [], 0, 0, tracer.w, tracer.h, name
        FakeHostGather.all.append(self)

    def begin_frame(self):
        self.calls.append("begin")

    def publish(self):
        self.calls.append("publish")
        self.kernels += 1
        self.frame_no += 1

    def wait_frame(self, keep=0):
        self.calls.append("wait%d" % keep)

    def frame(self, k):
        return np.zeros(self.w * self.h, np.uint32)

    def close(self):
        self.calls.append("close")


def run_bench(monkeypatch, argv, world=1):
    torch, dist = fake_torch()
    monkeypatch.setitem(sys.modules, "torch", torch)
    monkeypatch.setitem(sys.modules, "torch.distributed", dist)
    monkeypatch.setitem(sys.modules, "pynvml", fake_pynvml())
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    if world > 1:
        monkeypatch.setenv("WORLD_SIZE", str(world))
        monkeypatch.setenv("RANK", "0")
        monkeypatch.setenv("LOCAL_RANK", "0")
    import raytracer_rs_b200 as rt
    from raytracer_rs_b200 import multi_gpu

    FakeGather.all = []
    monkeypatch.setattr(multi_gpu, "FrameGather", FakeGather)
    FakeHostGather.all = []
    monkeypatch.setattr(multi_gpu, "HostFrameGather", FakeHostGather)

    tracers = []

    def from_scene(scene, cfg):
        tracers.append(FakeTracer(cfg.width, cfg.height, cfg.recursions))
        tracers[-1].cfg = cfg
        return tracers[-1]

    monkeypatch.setattr(rt.RayTracer, "from_scene", staticmethod(from_scene))
    spec = importlib.util.spec_from_file_location("bench_dryrun", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        bench.main()
    lines = [ln for ln in out.getvalue().splitlines() if ln.startswith("{")]
    assert len(lines) == 1  # ONE JSON line
    return json.loads(lines[0]), tracers


def test_bench_line_carries_the_whole_contract(monkeypatch):
    line, tracers = run_bench(monkeypatch, ["--steps", "6", "--warmup", "3", "--no-extras"])
    tracer = tracers[0]
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["metric"].startswith("Mrays/s") and line["unit"] == "Mrays/s" and line["n_gpus"] == 1
    assert (line["steps"], line["warmup"]) == (6, 3) and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["config"]["workload"] == "thai2_1080p" and "model" not in line["config"]
    assert line["config"]["primary_rays_per_step"] == PRIMARY and line["config"]["shadow_rays_per_step"] == SHADOW
    assert line["gpu_launches"] == 6  # one trace launch per timed step (the stand-in has no tile sort)
    assert line["clocks"]["sm_mhz"] == 1965 and line["clocks"]["reasons"] == []
    e2e = line["e2e"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(e2e) and e2e["d2h_bytes_per_step"] == 1920 * 1080 * 4 + 32
    assert e2e["h2d_bytes_per_step"] > 100 and e2e["frame_matches_device"] is True and e2e["sync_timeouts"] == 0
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["traffic_equivalent"] - roof["achieved"] / roof["peak"]) < 1e-12
    assert roof["algorithmic_bytes_per_launch"] == 6415965108.0 and roof["traffic"] > 10e6
    assert abs(roof["achieved"] - 6415965108.0 / 0.15e-3 / 1e9) < 1e-6 * roof["achieved"]
    issue = roof["issue_slots"]
    assert abs(issue["peak_ginst_s"] - 4 * 148 * 1.965) < 1e-9 and abs(issue["frac"] - issue["achieved_ginst_s"] / issue["peak_ginst_s"]) < 1e-12
    assert roof["frac"] == issue["frac"] and "issue" in roof["frac_of"]  # frac is the binding physical resource, not the traffic equivalent
    assert roof["stale"] in (True, False)  # whether the ncu counts were taken on the kernels the loaded library was built from
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["value"] > 0 and cpu["reference_work_matches_constants"] is True
    # the end-to-end leg hands every frame to the copy stream and waits for the previous one; camera state goes in every step
    assert tracer.calls.count("async") == 6 + 3 and tracer.calls.count("camera") == 6 + 3
    pipe = [c for c in tracer.calls if c in ("async", "wait1", "wait0")]  # frame i is queued before the host waits for frame i-1
    assert pipe == ["async", "wait1"] * 3 + ["wait0"] + ["async", "wait1"] * 6 + ["wait0"]
    assert tracer.tuning.get(10) == 0  # launch timing is off outside the roofline leg
    for key in ("value_long", "configs", "recursions2", "value_tree_walk", "e2e_reference_call_pattern", "first_frame_after_move", "strong"):
        assert key not in line


def test_bench_extras_on_one_gpu(monkeypatch):
    """The blocks the driver-run line carries beside the contract keys: a long run, the reference's band loop with incremental
    readback, the first frame after a camera move, the other BASELINE configurations (each with an oracle band check — the oracle
    really runs here, on the scene from the independent loader) and the RECURSIONS = 2 mode."""
    line, tracers = run_bench(monkeypatch, ["--steps", "3", "--warmup", "3", "--no-cpu-baseline"])
    assert "cpu_baseline" not in line
    long = line["value_long"]
    assert long["steps"] >= 3 and long["value"] > 0 and long["gpu_seconds"] > 0 and "sm_mhz" in long["clocks"]
    band = line["e2e_reference_call_pattern"]
    assert band["calls_per_frame"] == 22 and band["d2h_bytes_per_call"] == 50 * 1920 * 4 and band["d2h_bytes_per_call_full_frame"] == 1920 * 1080 * 4
    assert band["frame_matches_device"] is True and band["value"] > 0 and band["value_with_host_vec_clone"] > 0 and band["value_full_frame_readback"] > 0
    assert tracers[0].calls.count("delta") > 22 and tracers[0].calls.count("readback") > 22
    move = line["first_frame_after_move"]
    assert move["first_frame_after_move_ms"] > 0 and move["steady_state_ms"] > 0 and abs(move["ratio"] - move["first_frame_after_move_ms"] / move["steady_state_ms"]) < 1e-9
    assert tracers[0].calls.count("move") == 8
    cfgs = line["configs"]
    assert set(cfgs) == {"ico2_1024x768", "4boxes_1080p", "ico3_tex_1080p", "thai2_4k_16spp"}
    for name, c in cfgs.items():
        assert c["value"] > 0 and c["e2e"] > 0 and c["n_gpus"] == 1 and c["e2e_frame_matches_device"] is True
        chk = c["oracle_check"]  # the stand-in tracer renders nothing, so only the shape of the record is checked here
        assert set(chk) == {"band_rows", "ids_agree", "max_lsb_diff", "pixels_within_1_lsb"} and 0.0 <= chk["ids_agree"] <= 1.0
    assert cfgs["thai2_4k_16spp"]["spp"] == 16 and cfgs["thai2_4k_16spp"]["oracle_check"]["band_rows"] == [540, 556]
    rec = line["recursions2"]
    assert line["value_tree_walk"]["frame_and_film_equal_the_grid_path"] is True and line["value_tree_walk"]["value"] > 0
    assert rec["frame_ms"] > 0 and rec["rays_per_frame"]["bounce"] > 0 and rec["grays_per_s_all_rays"] > 0
    assert any(t.recursions == 2 for t in tracers)


def test_bench_launch_timing_flag(monkeypatch):
    line, tracers = run_bench(monkeypatch, ["--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--launch-timing", "--no-extras"])
    assert "cpu_baseline" not in line and abs(line["roofline"]["kernel_ms"] - 0.15) < 1e-9
    assert tracers[0].tuning.get(10) == 1


def test_bench_two_rank_control_flow(monkeypatch):
    """Rank 0 of a 2-GPU run (weak scaling: 2 samples per pixel, half the rows per rank): every frame is fenced, the
    device-timed leg releases the frame on the device, the end-to-end leg hands frame i to the copy stream BEFORE it
    waits for frame i-1, the last frame is waited for inside the timed region, and after the timed region every rank clears
    its film and ONE more gathered frame is compared with rank 0's unsharded render. The line also carries the strong-scaling
    block (the fixed 1-spp frame over the same ranks) and the 4K x 16 spp configuration sharded over the ranks."""
    line, tracers = run_bench(monkeypatch, ["--gpus", "2", "--steps", "4", "--warmup", "3"], world=2)
    assert line["n_gpus"] == 2 and line["scaling"] == "weak" and line["config"]["spp"] == 2 and "cpu_baseline" not in line
    assert line["roofline"]["traffic"] is None and "issue_slots" not in line["roofline"]
    assert line["gpu_launches"] == 4 + 4  # trace launch + fence launch per timed step
    assert line["e2e"]["frame_matches_device"] is True and line["e2e"]["sync_timeouts"] == 0
    calls = FakeGather.all[0].calls
    # device-timed legs (weak, long run, strong) fence every frame on the device; the NVLink gather is not used by the e2e legs here
    assert calls.count("gather_release") >= 2 * (4 + 3) and calls.count("gather") == 0 and calls.count("rearm") == 2
    hg = FakeHostGather.all[0].calls  # e2e leg of the weak configuration: every rank publishes its rows, rank 0 waits one frame behind
    assert hg == ["begin", "publish", "wait1"] * 3 + ["wait0"] + ["begin", "publish", "wait1"] * 4 + ["wait0"] + ["begin", "publish", "wait1", "wait0", "close"]
    assert len({g.name for g in FakeHostGather.all}) == len(FakeHostGather.all) == 3  # weak, strong, 4K x 16 spp: one shared segment each
    solo = [t for t in tracers if t.cfg.shard_count == 1 and t.cfg.width == 1920]
    assert len(solo) == 2 and all("close" in t.calls for t in solo)  # the unsharded check renders of the weak and the strong leg
    strong = line["strong"]
    assert strong["spp"] == 1 and strong["value"] > 0 and strong["e2e"]["frame_matches_device"] is True and strong["kernel_ms"] > 0
    assert ("configure", 0) in tracers[0].calls  # fixed jitter for the strong leg ...
    assert tracers[0].calls[-1] == "close" or ("configure", 1) in tracers[0].calls  # ... and back
    assert set(line["configs"]) == {"thai2_4k_16spp"} and line["configs"]["thai2_4k_16spp"]["n_gpus"] == 2
    assert "value_long" in line and "recursions2" not in line
