#!/usr/bin/env python
"""bench.py -- silhouette render+grad frames/s @1280x720, xArm7 mesh (BASELINE.json's metric).

One step = one pass of the hot path over one batch: B = 10 views of the xArm7 arm (links 1..7, 35,002
triangles), 1280x720, forward (per-link antialiased masks, sum, clamp, L2 loss against the reference masks)
+ backward (d loss / d mvp of every (view, link)).  value = frames/s with inputs resident in HBM;
e2e = the same through the host-buffer C-ABI call (H2D of the step's masks + matrices, D2H of loss and
gradient inside the timed region).  See DESIGN.md "Measurement" for the byte accounting.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, weak scaling: B views per rank)
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = dict(name="xarm7_links1-7_B10_1280x720_fwd+bwd", B=10, H=720, W=1280, links="xarm7", ring=4)
METRIC = "silhouette render+grad frames/sec @1280x720 xArm7 mesh"


def algorithmic_bytes_per_frame(H, W, V, F):
    """mask write 4HW + ref read 4HW + geometry fwd (12V+12F) + bwd (12V+12F) + grad-pos 16V."""
    return 8 * H * W + 40 * V + 24 * F


def build_sets(wl, rank, n_sets):
    """n_sets independent batches of B views: (meshes, [mvp (B,L,4,4) f32], link_poses) -- numpy, host side."""
    from easyhec_b200.scenes import make_scene, perturb_pose
    from util import scene_mvps
    sets = []
    for s in range(n_sets):
        sc = make_scene(wl["B"], wl["H"], wl["W"], links=wl["links"], seed=1000 * rank + s)
        rng = np.random.RandomState(77 + 1000 * rank + s)
        mvp_gt = scene_mvps(sc, wl["H"], wl["W"])
        mvp = scene_mvps(sc, wl["H"], wl["W"], perturb_pose(sc["Tc_c2b"], rng, 0.03, 3.0))
        sets.append(dict(scene=sc, mvp_gt=mvp_gt, mvp=mvp))
    return sets


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(wl, n_views, steps, warmup, threads=None):
    """Times oracle.render_views (the CPU restatement of the path) on n_views views of the workload.
    Returns (frames_per_s, ms_per_step, threads)."""
    from oracle import oracle
    from util import scene_mvps  # noqa: F401
    from easyhec_b200.scenes import make_scene, perturb_pose
    B = wl["B"]
    nsets = (n_views + B - 1) // B
    sets = build_sets(wl, 0, nsets)
    packed = oracle.pack_links(sets[0]["scene"]["meshes"])
    mvp = np.concatenate([s["mvp"] for s in sets])[:n_views]
    mvp_gt = np.concatenate([s["mvp_gt"] for s in sets])[:n_views]
    ref = oracle.union_binary(packed, mvp_gt, wl["H"], wl["W"]).astype(np.float32)
    for _ in range(warmup):
        oracle.render_views(packed, mvp, ref, wl["H"], wl["W"])
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.render_views(packed, mvp, ref, wl["H"], wl["W"])
    dt = time.perf_counter() - t0
    return n_views * steps / dt, 1e3 * dt / max(steps, 1), oracle.num_threads()


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.  nvdiffrast (the reference's GPU
    arithmetic) is not vendored in the reference tree and cannot be installed offline, so this arm times the
    oracle port (kind "port") with all host threads; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOAD
    from oracle import oracle
    oracle.build()
    threads = oracle.num_threads()
    # calibrate: one view-step
    t_one, _, _ = cpu_reference_run(wl, max(1, min(threads, 4)), 1, 0)
    budget = 150.0
    total_steps = args.steps + args.warmup
    views = int(max(1, min(4 * threads, budget * t_one / max(total_steps, 1))))
    fps, ms, threads = cpu_reference_run(wl, views, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "views_per_step": views, "H": wl["H"], "W": wl["W"]},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": "%d views of the workload per step, OpenMP over views" % views},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference's own GPU arithmetic (nvdiffrast) is not in the reference tree / not installable "
                    "offline; this is the CPU restatement in oracle/ (parity unpinned, see DESIGN.md)"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
class JsonStdout:
    """Under torchrun, libraries print to stdout while they initialise (NCCL prints its version from rank 0).  The
    contract is ONE JSON line on stdout: file descriptor 1 is pointed at stderr for the duration of the run and the
    line is written to the saved descriptor at the end."""

    def __init__(self, active):
        self.fd = None
        if active:
            sys.stdout.flush()
            self.fd = os.dup(1)
            os.dup2(2, 1)

    def emit(self, obj):
        data = (json.dumps(obj) + "\n").encode()
        sys.stdout.flush()
        if self.fd is None:
            sys.stdout.write(data.decode())
            sys.stdout.flush()
        else:
            os.write(self.fd, data)


def run_b200(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = JsonStdout(world > 1)
    import torch
    import torch.distributed as dist
    from easyhec_b200._lib import Context

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOAD
    B, H, W, R = wl["B"], wl["H"], wl["W"], wl["ring"]
    ctx = Context(dev)
    if os.environ.get("EHB_PIPES"):
        ctx.set_pipelines(int(os.environ["EHB_PIPES"]))
    sets = build_sets(wl, rank, R)
    meshes = sets[0]["scene"]["meshes"]
    ids = [ctx.register_mesh(m.vertices, m.faces) for m in meshes]
    L = len(ids)
    V = sum(len(m.vertices) for m in meshes)
    F = sum(len(m.faces) for m in meshes)
    ctx.reserve(B, L, F, H, W)
    # device-resident inputs: reference masks rendered by the binary path at the ground-truth pose
    mvp_dev = [torch.from_numpy(s["mvp"]).to(dev) for s in sets]
    ref_dev = []
    for s in sets:
        m = ctx.render_binary_batch(ids, torch.from_numpy(s["mvp_gt"]).to(dev), H, W)
        ref_dev.append(m.to(torch.float32))
    masks = [torch.empty((B, H, W), dtype=torch.float32, device=dev) for _ in range(R)]
    loss = torch.empty((B,), dtype=torch.float64, device=dev)
    gmvp = torch.empty((B, L, 4, 4), dtype=torch.float64, device=dev)
    g7 = torch.zeros(7, dtype=torch.float32, device=dev)
    collective = "none"
    if world > 1:
        collective = "nccl all-reduce 7xf32 per step"
        if args.collective == "peer":
            try:   # one-shot all-reduce through NVLink peer mailboxes (ehb_allreduce7); NCCL stays the fallback
                ctx.comm_connect()
                collective = "NVLink peer-mailbox all-reduce 7xf32 per step (ehb_allreduce7)"
            except Exception as e:   # noqa: BLE001
                if rank == 0:
                    print("peer all-reduce unavailable (%s); using NCCL" % e, file=sys.stderr)
        agree = torch.tensor([1.0 if collective.startswith("NVLink") else 0.0], device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if agree.item() < 1.0:
            collective = "nccl all-reduce 7xf32 per step"
    use_peer = collective.startswith("NVLink")

    def step(k):
        s = k % R
        ctx.render_views_fused(ids, mvp_dev[s], ref_dev[s], H, W, backward=True, out=(masks[s], loss, gmvp))
        if world > 1:   # the solver's one exchange step: all-reduce of (g_dof[6], loss) -- trainer/base.py:349
            if use_peer:
                ctx.allreduce7(g7)
            else:
                dist.all_reduce(g7)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(max(args.warmup, 3)):
        step(k)
    flags, nclip = ctx.status()
    assert flags & 1 == 0, "scratch overflow during warm-up"
    # The step is a fixed launch sequence (pipelines forked / joined with events): capture one CUDA graph per ring slot and
    # replay them, so that the timed loop is not bound by the host's launch rate.  EHB_BENCH_NOGRAPH=1 keeps eager launches.
    graphs, launches_per_step = None, None
    if not os.environ.get("EHB_BENCH_NOGRAPH") and (world == 1 or use_peer):
        try:
            cap = torch.cuda.Stream()
            cap.wait_stream(torch.cuda.current_stream())
            graphs = []
            with torch.cuda.stream(cap):
                step(0)
                torch.cuda.synchronize()
                for s in range(R):
                    g = torch.cuda.CUDAGraph()
                    l_before = ctx.launch_count()
                    with torch.cuda.graph(g, stream=cap):
                        step(s)
                    launches_per_step = ctx.launch_count() - l_before
                    graphs.append(g)
            torch.cuda.current_stream().wait_stream(cap)
            torch.cuda.synchronize()
        except Exception as e:   # noqa: BLE001
            graphs = None
            if rank == 0:
                print("CUDA graph capture failed (%s); eager launches" % str(e)[:200], file=sys.stderr)
    eager_step = step
    if graphs is not None:
        def step(k):   # noqa: F811
            graphs[k % R].replay()
        for k in range(2 * R):
            step(k)
    # ---- timed region (device-resident inputs) ------------------------------------------------------------
    sampler = ClockSampler(physical_gpu_index(local))
    barrier()
    sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        step(k)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    if graphs is not None:
        launches = launches_per_step * args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    frames = B * args.steps * world
    value = frames / (ms * 1e-3)

    # ---- per-kernel pass (CUDA events around each kernel, on the launching stream) ----------------------
    ctx.profile(True)
    ctx.kernel_times()
    nprof = min(args.steps, 200)
    for k in range(nprof):
        eager_step(k)
    kt, npass = ctx.kernel_times()
    ctx.profile(False)
    kavg_us = {k: 1e3 * v / max(npass, 1) for k, v in kt.items()}
    dom = max(kavg_us, key=kavg_us.get)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = algorithmic_bytes_per_frame(H, W, V, F) * B
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    achieved = alg / (kavg_us[dom] * 1e-6) / 1e9
    roofline = {"bound": "hbm", "kernel": "ehb_k_" + dom + ("(+ehb_k_raster_big)" if dom == "raster" else ""), "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "algorithmic_bytes_per_launch": alg, "kernel_us": kavg_us,
                "kernel_us_note": "CUDA events around each stage, single pipeline; the timed region overlaps pipelines",
                "step_frac": (alg / (sum(kavg_us.values()) * 1e-6) / 1e9) / peak}

    # ---- end to end: host buffers through the C ABI, copies inside the timed region ----------------------
    # every step: H2D of that step's reference masks (u8, what the dataset holds before .float()) and matrices from
    # pinned host memory, the fused pass, D2H of loss + gradient; up to four steps in flight so that one step's copies
    # overlap the others' kernels and the PCIe link never waits for the host.  Each step's result is complete on the host when its _end returns.
    mvp_host = [torch.from_numpy(s["mvp"]).pin_memory() for s in sets]
    ref_host = [r.to(torch.uint8).cpu().pin_memory() for r in ref_dev]
    S = max(2, min(4, int(os.environ.get("EHB_E2E_SLOTS", "4"))))   # steps in flight: the next H2D is always queued
    loss_host = [torch.empty((B,), dtype=torch.float64).pin_memory() for _ in range(S)]
    gmvp_host = [torch.empty((B, L, 4, 4), dtype=torch.float64).pin_memory() for _ in range(S)]

    def e2e_run(n):
        for k in range(n):
            ctx.solver_step_begin_u8(k % S, ids, mvp_host[k % R], ref_host[k % R], H, W, loss_host[k % S], gmvp_host[k % S])
            if k >= S - 1:
                ctx.solver_step_end((k - S + 1) % S)
        for k in range(max(n - S + 1, 0), n):
            ctx.solver_step_end(k % S)

    e2e_run(16)
    e2e_steps = min(args.steps, 1000)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    ms_e2e = 1e3 * (time.perf_counter() - t0)   # host clock: every step ends with a host-visible result
    barrier()
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e = {"value": B * e2e_steps * world / (ms_e2e * 1e-3), "unit": "frames/s",
           "h2d_bytes_per_step": int(B * H * W + B * L * 64), "d2h_bytes_per_step": int(8 * B + 128 * B * L + 176),
           "steps": e2e_steps, "api": "ehb_solver_step_begin_u8 / _end, %d slots (pinned host masks u8 + mvp in, " % S +
                                      "loss + g_mvp out, host-visible result every step)"}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle import oracle
            oracle.build()
            nv = max(1, min(oracle.num_threads(), 16))
            fps, cms, threads = cpu_reference_run(wl, nv, 2, 1)
            cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": "%d views of the workload x 2 steps (+1 warm-up), OpenMP over views" % nv}
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "views_per_step_per_gpu": B, "H": H, "W": W, "links": L,
                           "triangles": F, "vertices": V,
                           "l2": "ring of %d view-sets (%.0f MB of masks+refs) > 126 MB L2" %
                                 (R, R * B * H * W * 8 / 1e6),
                           "pipelines": int(os.environ.get("EHB_PIPES", "3")),
                           "launch": "CUDA graph replay, one graph per ring slot" if graphs is not None else "eager",
                           "collective": collective},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "need_clip_triangles": int(nclip)}
        out.emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"], help="N>1: how the 7 floats are all-reduced")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
