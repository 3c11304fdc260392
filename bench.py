#!/usr/bin/env python
"""bench.py -- silhouette render+grad frames/s @1280x720, xArm7 mesh (BASELINE.json's metric).

One step = one pass of the hot path over one batch: B = 10 views of the xArm7 arm (links 1..7, 35,002
triangles), 1280x720, forward (per-link antialiased masks, sum, clamp, L2 loss against the reference masks)
+ backward to the 6-DoF pose gradient (d loss / d mvp of every (view, link), chained to d loss / d dof) --
SURVEY.md 8(d): "one frame = one view, all links, forward + backward to g_dof".
value = frames/s with inputs resident in HBM; e2e = the same through the host-buffer C-ABI call (host -> device
copy of the step's inputs, device -> host copy of loss and gradient inside the timed region).
See DESIGN.md "Measurement" for the byte accounting.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload headline|inview|explore]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, weak scaling: B views per rank)
"""
import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# headline: the reference renderer's own demo pose (nvdiffrast_renderer.py:77-80) with random joints -- about half of
#           the arm's vertices project off screen.  inview: same robot, joints and intrinsics, the camera translated so
#           that >= 95 % of the vertices are on screen (scenes.fit_camera): twice the triangles to draw.
WORKLOADS = {
    "headline": dict(name="xarm7_links1-7_B10_1280x720_fwd+bwd", B=10, H=720, W=1280, links="xarm7", ring=4, camera="sample"),
    "inview": dict(name="xarm7_links1-7_B10_1280x720_fwd+bwd_inview", B=10, H=720, W=1280, links="xarm7", ring=4, camera="fit"),
}
WORKLOAD = WORKLOADS["headline"]
EXPLORE = dict(name="explore_xarm7_all_Q256xC4_1920x1080", Q=256, C=4, H=1080, W=1920)   # BASELINE.json configs[3]
METRIC = "silhouette render+grad frames/sec @1280x720 xArm7 mesh"
MIN_TIMED_MS = 200.0   # the K-step timed region is repeated until this much device time has been measured


def algorithmic_bytes_per_frame(H, W, V, F):
    """mask write 4HW + ref read 4HW + geometry fwd (12V+12F) + bwd (12V+12F) + grad-pos 16V."""
    return 8 * H * W + 40 * V + 24 * F


def algorithmic_bytes_per_render(H, W, V, F):
    """SURVEY.md 8(d), forward-only binary render: 4HW + 12V + 12F."""
    return 4 * H * W + 12 * V + 12 * F


def build_sets(wl, rank, n_sets, same_on_every_rank=False):
    """n_sets independent batches of B views: dicts(scene, mvp_gt, mvp (B,L,4,4) f32, Tc (4,4) perturbed pose) -- numpy."""
    from easyhec_b200.scenes import fit_camera, make_scene, perturb_pose
    from util import scene_mvps
    sets = []
    r = 0 if same_on_every_rank else rank
    for s in range(n_sets):
        sc = make_scene(wl["B"], wl["H"], wl["W"], links=wl["links"], seed=1000 * r + s)
        if wl.get("camera") == "fit":
            sc["Tc_c2b"] = fit_camera(sc, wl["H"], wl["W"], keep=0.96)
        rng = np.random.RandomState(77 + 1000 * r + s)
        mvp_gt = scene_mvps(sc, wl["H"], wl["W"])
        Tc = perturb_pose(sc["Tc_c2b"], rng, 0.03, 3.0)
        mvp = scene_mvps(sc, wl["H"], wl["W"], Tc)
        sets.append(dict(scene=sc, mvp_gt=mvp_gt, mvp=mvp, Tc=Tc))
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
def host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: not inherited here)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_reference_run(wl, sets, steps, warmup):
    """Times oracle.render_views (the CPU restatement of the path) on the views of `sets`, one set per step in turn, with
    every host thread.  Returns (frames_per_s, ms_per_step, threads, results of the last pass over every set)."""
    from oracle import oracle
    threads = oracle.set_num_threads(host_threads())
    packed = oracle.pack_links(sets[0]["scene"]["meshes"])
    refs = [oracle.union_binary(packed, s["mvp_gt"], wl["H"], wl["W"]).astype(np.float32) for s in sets]
    last = [None] * len(sets)
    for k in range(warmup):
        oracle.render_views(packed, sets[k % len(sets)]["mvp"], refs[k % len(sets)], wl["H"], wl["W"])
    t0 = time.perf_counter()
    for k in range(steps):
        i = k % len(sets)
        last[i] = oracle.render_views(packed, sets[i]["mvp"], refs[i], wl["H"], wl["W"])
    dt = time.perf_counter() - t0
    return wl["B"] * steps / dt, 1e3 * dt / max(steps, 1), threads, last


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores.  nvdiffrast (the reference's GPU
    arithmetic) is not vendored in the reference tree and cannot be installed offline, so this arm times the
    oracle port (kind "port") with all host threads on the GPU arm's own config: every step is the same B views."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS.get(args.workload, WORKLOAD)
    from oracle import oracle
    oracle.build()
    sets = build_sets(wl, 0, wl["ring"])
    # bounded: the whole run (warm-up + K steps) stays within a few minutes on any host
    t_probe = cpu_reference_run(wl, sets[:1], 1, 0)[1] * 1e-3
    steps = int(max(1, min(args.steps, 150.0 / max(t_probe, 1e-3))))
    warm = int(min(args.warmup, max(0, 30.0 / max(t_probe, 1e-3))))
    fps, ms, threads, _ = cpu_reference_run(wl, sets, steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "views_per_step_per_gpu": wl["B"], "H": wl["H"], "W": wl["W"],
                       "steps_timed": steps},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": "%d views of the workload per step (the GPU arm's batch), %d steps, OpenMP over "
                                       "views, %d threads set explicitly" % (wl["B"], steps, threads)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference's own GPU arithmetic (nvdiffrast) is not in the reference tree / not installable "
                    "offline; this is the CPU restatement in oracle/ (parity unpinned, see DESIGN.md)"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- baselines beside the GPU arm
def time_proxy(wl, sets, ids_unused, dev, iters=3):
    """Structural proxy of the reference's call pattern -- NOT nvdiffrast: RBSolver(fused=False) calls the drop-in
    operator once per (view, link) (rb_solver.py:60-67), sums / clamps / takes the loss with torch ops and lets torch
    autograd drive one operator backward per call.  Same kernels as the fused path, none of its fusion."""
    import torch
    from easyhec_b200.rb_solver import RBSolver
    s = sets[0]
    sc = s["scene"]
    m = RBSolver(meshes=sc["meshes"], init_Tc_c2b=s["Tc"], H=wl["H"], W=wl["W"], fused=False, device=dev,
                 want_outputs=False)
    ctx = m.renderer.ctx
    ref = ctx.render_binary_batch(m.mesh_ids, torch.from_numpy(s["mvp_gt"]).to(dev), wl["H"], wl["W"]).float()
    dps = {"mask": ref, "link_poses": torch.from_numpy(sc["link_poses"]).to(dev), "K": torch.from_numpy(sc["K"]).to(dev)[None],
           "global_step": 0}

    def one():
        m.zero_grad(set_to_none=True)
        _, ld = m(dps)
        ld["mask_loss"].backward()

    one()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(iters):
        one()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    ctx.close()
    return {"kind": "unfused per-(view, link) operator loop over this repo's own kernels + torch autograd -- NOT nvdiffrast",
            "value": wl["B"] * iters / dt, "unit": "frames/s", "steps": iters, "calls_per_step": 2 * wl["B"] * len(sc["meshes"])}


def time_nvdiffrast(wl, sets, dev, iters=3):
    """The reference's own GPU path (nvdiffrast_renderer.py:25-48 driven as rb_solver.py:60-72), timed iff nvdiffrast
    imports on this box."""
    try:
        import nvdiffrast.torch as dr
    except Exception as e:   # noqa: BLE001
        return "unavailable: %s" % type(e).__name__
    import torch
    from easyhec_b200.projection import K_to_projection, opencv2gl
    s = sets[0]
    sc = s["scene"]
    H, W = wl["H"], wl["W"]
    glctx = dr.RasterizeCudaContext()
    verts = [torch.from_numpy(np.asarray(m.vertices, np.float32)).to(dev) for m in sc["meshes"]]
    faces = [torch.from_numpy(np.asarray(m.faces, np.int32)).to(dev) for m in sc["meshes"]]
    K = torch.from_numpy(sc["K"]).to(dev)
    lp = torch.from_numpy(sc["link_poses"]).to(dev)
    Tc = torch.tensor(s["Tc"], dtype=torch.float32, device=dev, requires_grad=True)
    flip = opencv2gl(dev)
    ref = torch.zeros((wl["B"], H, W), device=dev)

    def render_mask(v, f, pose):
        proj = K_to_projection(K, H, W).to(dev)
        pos = torch.cat([v, torch.ones_like(v[:, :1])], 1) @ (proj @ (flip @ pose)).T
        rast, _ = dr.rasterize(glctx, pos[None], f, resolution=[H, W])
        col, _ = dr.interpolate(torch.ones_like(v)[None], rast, f)
        col = dr.antialias(col, rast, pos[None], f)
        return torch.flip(col[0, :, :, 0], dims=[0])

    def one():
        losses = []
        for b in range(wl["B"]):
            si = torch.stack([render_mask(verts[l], faces[l], Tc @ lp[b, l]) for l in range(len(verts))]).sum(0).clamp(max=1)
            losses.append(torch.sum((si - ref[b]) ** 2))
        torch.stack(losses).mean().backward()

    one()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(iters):
        one()
    torch.cuda.synchronize(dev)
    return {"value": wl["B"] * iters / (time.perf_counter() - t0), "unit": "frames/s", "steps": iters}


def probe_pytorch3d():
    try:
        import pytorch3d  # noqa: F401
        return "importable but not timed (CPU soft rasterizer of north_star; see DESIGN.md)"
    except Exception as e:   # noqa: BLE001
        return "unavailable: %s" % type(e).__name__


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


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def graph_steps(k, cap, unit):
    """Steps per captured graph: the largest multiple of `unit` that is <= cap and divides k; without such a divisor the
    largest multiple of `unit` <= min(cap, k) (at least `unit`)."""
    best = 0
    for g in range(unit, min(cap, k) + 1, unit):
        if k % g == 0:
            best = g
    return best or max(unit, min(cap, k) - min(cap, k) % unit)


def timed_regions(run, steps, barrier, torch, min_ms=MIN_TIMED_MS, max_repeats=400, agree=None):
    """Times EXACTLY `steps` steps (run(steps, first_step_index)) between two CUDA events, bracketed by barrier +
    synchronize on both sides; the region is repeated until `min_ms` of device time has been measured (a 20-step region
    of a 0.1 ms step is 2 ms: one region is noise).  Returns (median ms per region, all region times)."""
    regions = []
    k0 = 0
    while True:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps, k0)
        e1.record()
        barrier()
        regions.append(e0.elapsed_time(e1))
        k0 += steps
        done = sum(regions) >= min_ms or len(regions) >= max_repeats
        if agree is not None:      # several ranks: every rank must run the same number of regions (each one ends in a barrier)
            done = agree(done)
        if done:
            break
    return float(np.median(regions)), regions


def run_b200(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = JsonStdout(world > 1)
    import torch
    import torch.distributed as dist
    from easyhec_b200._lib import Context
    from easyhec_b200.scenes import onscreen_fraction
    from easyhec_b200.se3 import matrix_to_dof

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    B, H, W, R = wl["B"], wl["H"], wl["W"], wl["ring"]
    ctx = Context(dev)
    if os.environ.get("EHB_PIPES"):
        ctx.set_pipelines(int(os.environ["EHB_PIPES"]))
    # default: the library splits a call over 3 pipelines only when it has more than 16 items
    pipes_used = int(os.environ["EHB_PIPES"]) if os.environ.get("EHB_PIPES") else (1 if wl["B"] <= 16 else 3)
    # every rank renders the SAME scenes (weak scaling: B views per rank): the ranks' work is balanced by construction, so
    # what the N-GPU line shows is the cost of the exchange, not of unequal scenes
    sets = build_sets(wl, rank, R, same_on_every_rank=True)
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
    # the reference masks are constant over a solve: registered once with the context (bit-packed + per-tile counts)
    ref_h = [ctx.register_ref(r.to(torch.uint8)) for r in ref_dev]
    coverage = float(torch.stack([r.mean() for r in ref_dev]).mean().item())
    onscreen = float(np.mean([onscreen_fraction(s["scene"], H, W) for s in sets]))
    masks = [torch.empty((B, H, W), dtype=torch.float32, device=dev) for _ in range(R)]
    loss = torch.empty((B,), dtype=torch.float64, device=dev)
    gmvp = torch.empty((B, L, 4, 4), dtype=torch.float64, device=dev)
    # the pose chain behind the rasterizer: d loss/d mvp -> d loss/d dof (+ loss) = the 7 floats a data-parallel run exchanges
    dof_dev = [matrix_to_dof(torch.tensor(s["Tc"], dtype=torch.float32)).to(dev).contiguous() for s in sets]
    K_dev = torch.from_numpy(sets[0]["scene"]["K"]).to(dev).contiguous()
    lp_dev = [torch.from_numpy(s["scene"]["link_poses"]).to(dev).contiguous() for s in sets]
    g7 = torch.zeros(7, dtype=torch.float32, device=dev)
    dof_scratch = dof_dev[0].clone()       # Adam runs on a scratch copy: the rendered poses stay those of the ring
    adam_state = torch.zeros(13, dtype=torch.float32, device=dev)
    collective = "none"
    if world > 1:
        collective = "nccl all-reduce 7xf32 per step"
        if args.collective == "peer":
            try:   # one-shot all-reduce through NVLink peer mailboxes (ehb_allreduce7); NCCL stays the fallback
                ctx.comm_connect()
                collective = "NVLink peer-mailbox all-reduce 7xf32 per step, fused into pose_backward (send) and adam (recv)"
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
        ctx.render_views_fused(ids, mvp_dev[s], ref_h[s], H, W, backward=True, out=(masks[s], loss, gmvp))
        # the solver's one exchange step: all-reduce of (g_dof[6], loss) -- trainer/base.py:349 -- then Adam.  One rank runs
        # the same chain without the exchange, so that the N-GPU lines differ from the 1-GPU line by the exchange alone.
        if world == 1 or use_peer:   # pose chain (+ send), (recv +) Adam in one launch
            ctx.pose_backward_adam(dof_dev[s], K_dev, lp_dev[s], gmvp, loss, H, W, adam_state, 3e-3, weight_decay=5e-4,
                                   grad_scale=1.0 / world, loss_scale=1.0 / (B * world), out=g7, exchange=use_peer,
                                   adam_dof=dof_scratch)
        else:
            ctx.pose_backward(dof_dev[s], K_dev, lp_dev[s], gmvp, loss, H, W, grad_scale=1.0 / world,
                              loss_scale=1.0 / (B * world), out=g7)
            dist.all_reduce(g7)
            ctx.adam_step(dof_scratch, g7, adam_state, 3e-3, weight_decay=5e-4)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def agree(done):               # a region loop ends when EVERY rank has measured enough
        if world == 1:
            return done
        t = torch.tensor([1.0 if done else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() >= 1.0)

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
    def run_serial(n, k0):
        for k in range(n):
            step(k0 + k)

    # ---- steps in flight: the same step submitted to the context's four slots (ehb_step_begin), each slot a stream with its
    # own scratch -- independent batches (the views of different solves, exploration rounds, the ring slots here) overlap
    # one step's latency-bound kernels and tails with the others' work.  Step k runs on slot k % S on every rank, so the
    # exchanges of a slot (its own mailbox channel) pair up across the ranks.
    S = max(1, min(8, int(os.environ.get("EHB_VALUE_SLOTS", "4"))))
    inflight = S > 1 and (world == 1 or use_peer)
    # steps per captured graph (the slots drain at a graph's end): as many as 128, a multiple of the ring and the slots, and
    # when possible a divisor of K, so that a timed region is graph replays only (a remainder runs as eager launches)
    G = int(os.environ["EHB_VALUE_GRAPH"]) if os.environ.get("EHB_VALUE_GRAPH") else graph_steps(args.steps, 128, R * S // math.gcd(R, S))
    g7s = [torch.zeros(7, dtype=torch.float32, device=dev) for _ in range(S)]
    dof_scr = [dof_dev[0].clone() for _ in range(S)]
    adam_st = [torch.zeros(13, dtype=torch.float32, device=dev) for _ in range(S)]

    def slot_step(k):
        s_, sl = k % R, k % S
        ctx.step_begin(sl, ids, ref_h[s_], H, W, mvp_dev[s_], masks=masks[s_], dof=dof_dev[s_], K=K_dev, link_poses=lp_dev[s_],
                       out7=g7s[sl], adam_dof=dof_scr[sl], adam_state=adam_st[sl],
                       lr=3e-3, weight_decay=5e-4, grad_scale=1.0 / world, loss_scale=1.0 / (B * world), exchange=use_peer)

    slot_graph, slot_launches = None, None
    if inflight:
        ctx.slots_fork()
        for k in range(2 * S):
            slot_step(k)
        ctx.slots_join()
        torch.cuda.synchronize()
        for sl in range(S):
            ctx.solver_step_end(sl)            # raises if a slot's scratch overflowed
        if not os.environ.get("EHB_BENCH_NOGRAPH"):
            try:
                cap = torch.cuda.Stream()
                cap.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(cap):
                    slot_graph = torch.cuda.CUDAGraph()
                    l_before = ctx.launch_count()
                    with torch.cuda.graph(slot_graph, stream=cap):
                        ctx.slots_fork()
                        for k in range(G):
                            slot_step(k)
                        ctx.slots_join()
                    slot_launches = (ctx.launch_count() - l_before) / G
                torch.cuda.current_stream().wait_stream(cap)
                torch.cuda.synchronize()
            except Exception as e:   # noqa: BLE001
                slot_graph = None
                if rank == 0:
                    print("CUDA graph capture of the slot steps failed (%s); eager" % str(e)[:200], file=sys.stderr)

    def run_inflight(n, k0):
        q, r_ = (n // G, n % G) if slot_graph is not None else (0, n)
        for _ in range(q):
            slot_graph.replay()
        if r_:
            ctx.slots_fork()
            for k in range(r_):
                slot_step(k0 + q * G + k)      # (G is a multiple of the ring and of S: the slot / ring pairing continues)
            ctx.slots_join()

    # ---- timed regions (device-resident inputs) ------------------------------------------------------------
    sampler = ClockSampler(physical_gpu_index(local))
    barrier()
    sampler.start()
    l0 = ctx.launch_count()
    ms_serial, regions_serial = timed_regions(run_serial, args.steps, barrier, torch, agree=agree)
    launches = (ctx.launch_count() - l0) // max(len(regions_serial), 1)
    if graphs is not None:
        launches = launches_per_step * args.steps
    ms, regions = ms_serial, regions_serial
    if inflight:
        for k in range(2):
            run_inflight(G, 0)
        l0 = ctx.launch_count()
        ms, regions = timed_regions(run_inflight, args.steps, barrier, torch, agree=agree)
        launches = (ctx.launch_count() - l0) // max(len(regions), 1)
        if slot_graph is not None:
            launches = int(round(slot_launches * (args.steps - args.steps % G))) + (launches if args.steps % G else 0)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms, ms_serial], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_serial = float(t[0].item()), float(t[1].item())
    frames = B * args.steps * world
    value = frames / (ms * 1e-3)
    serial = {"value": frames / (ms_serial * 1e-3), "unit": "frames/s", "ms_per_step": ms_serial / max(args.steps, 1),
              "launch": "CUDA graph replay, one graph per ring slot" if graphs is not None else "eager",
              "what": "one step at a time on one stream (each step's views split over %d pipelines): the latency of a step, "
                      "what a single sequential solve sees" % pipes_used}

    # ---- parity of what was just timed: GPU results of every ring slot, checked by the CPU leg below -------------
    gpu_results, slot_path_ok = [], None
    if rank == 0 and world == 1 and not args.no_cpu:
        for s in range(R):
            eager_step(s)
            gpu_results.append((masks[s].cpu().numpy(), loss.cpu().numpy().copy(), gmvp.cpu().numpy().copy(), g7.cpu().numpy().copy()))
        if inflight:   # the slot path on the same inputs: masks and the 7 floats must equal the single-stream path's
            slot_path_ok = True
            for s in range(R):
                masks[s].fill_(-1.0)
            ctx.slots_fork()
            for s in range(R):
                slot_step(s)
            ctx.slots_join()
            torch.cuda.synchronize()
            for s in range(R):
                slot_path_ok &= bool(np.array_equal(masks[s].cpu().numpy(), gpu_results[s][0]))
                slot_path_ok &= bool(np.allclose(g7s[s % S].cpu().numpy(), gpu_results[s][3], rtol=1e-5, atol=1e-7))

    # ---- per-stage pass: CUDA events around each stage on the launching stream, single pipeline.  The passes are captured
    # (one graph per ring slot, the events become event-record nodes) and replayed, so that a stage is timed with a
    # graph's launch gaps, not with the eager launch overhead of four kernels and five event records ----------------
    ctx.profile(True)
    ctx.kernel_times()
    nprof = min(max(args.steps, 50), 200)
    kt, npass, stage_launch = {"front": 0.0, "raster": 0.0, "tiles": 0.0}, 0, "eager"
    pgraphs = None
    if not os.environ.get("EHB_BENCH_NOGRAPH"):
        try:
            cap = torch.cuda.Stream()
            cap.wait_stream(torch.cuda.current_stream())
            pgraphs = []
            with torch.cuda.stream(cap):
                for s in range(R):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=cap):
                        eager_step(s)
                    pgraphs.append(g)
            torch.cuda.current_stream().wait_stream(cap)
            torch.cuda.synchronize()
        except Exception as e:   # noqa: BLE001
            pgraphs = None
            ctx.kernel_times()
            if rank == 0:
                print("capture of the profiled pass failed (%s); eager stage timing" % str(e)[:200], file=sys.stderr)
    if pgraphs is not None:
        stage_launch = "CUDA graph replay"
        for k in range(2 * R):
            pgraphs[k % R].replay()
        for _ in range(max(nprof // R, 1)):
            for g in pgraphs:
                g.replay()
            t_, n_ = ctx.kernel_times(peek=True)      # (synchronises) the times of the R passes just replayed
            for k_ in kt:
                kt[k_] += t_[k_]
            npass += n_
        ctx.kernel_times()
    else:
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
                "traffic_source": "ncu --set full capture of the same command, committed as profiles/traffic.json",
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "algorithmic_bytes_per_launch": alg, "kernel_us": kavg_us,
                "kernel_us_note": "CUDA events around each stage (front = table | vertices | batch lists | clear | tile lists; "
                                  "raster = k_raster + k_raster_big; tiles = image-space stage), single pipeline, %s; "
                                  "the timed region overlaps steps" % stage_launch,
                "step_frac": (alg / (sum(kavg_us.values()) * 1e-6) / 1e9) / peak,
                "timed_step_frac": (alg / (ms / max(args.steps, 1) * 1e-3) / 1e9) / peak}

    # ---- end to end: host buffers through the C ABI, copies inside the timed region ----------------------
    # every step: H2D of that step's inputs (the matrices of its B x L (view, link) pairs, from pinned host memory -- the
    # reference masks were registered once, they are not an input of a step), the fused pass, D2H of loss + gradient; up to
    # four steps in flight so that one step's copies overlap the others' kernels.  Each step's result is complete on the
    # host when its _end returns.
    mvp_host = [torch.from_numpy(s["mvp"]).pin_memory() for s in sets]
    Se = max(2, min(8, int(os.environ.get("EHB_E2E_SLOTS", "4"))))   # steps in flight
    loss_host = [torch.empty((B,), dtype=torch.float64).pin_memory() for _ in range(Se)]
    gmvp_host = [torch.empty((B, L, 4, 4), dtype=torch.float64).pin_memory() for _ in range(Se)]

    def e2e_run(n, begin):
        for k in range(n):
            begin(k)
            if k >= Se - 1:
                ctx.solver_step_end((k - Se + 1) % Se)
        for k in range(max(n - Se + 1, 0), n):
            ctx.solver_step_end(k % Se)

    def begin_ref(k):
        ctx.solver_step_begin_ref(k % Se, ids, mvp_host[k % R], ref_h[k % R], H, W, loss_host[k % Se], gmvp_host[k % Se])

    def e2e_time(begin, n):
        e2e_run(16, begin)
        barrier()
        t0 = time.perf_counter()
        e2e_run(n, begin)
        dt = 1e3 * (time.perf_counter() - t0)   # host clock: every step ends with a host-visible result
        barrier()
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    e2e_steps = min(max(args.steps, 200), 2000)
    ms_e2e = e2e_time(begin_ref, e2e_steps)
    e2e = {"value": B * e2e_steps * world / (ms_e2e * 1e-3), "unit": "frames/s",
           "h2d_bytes_per_step": int(B * L * 64), "d2h_bytes_per_step": int(8 * B + 128 * B * L + 176),
           "steps": e2e_steps, "api": "ehb_solver_step_begin_ref / _end, %d slots (pinned host mvp in, loss + g_mvp out, " % Se +
                                      "host-visible result every step; reference masks registered once with ehb_ref_register)"}
    # the reference trainer's quirk, for comparison: the masks re-uploaded on every step (trainer/rbsolver.py:31 to_cuda(batch))
    ref_host = [r.to(torch.uint8).cpu().pin_memory() for r in ref_dev]

    def begin_u8(k):
        ctx.solver_step_begin_u8(k % Se, ids, mvp_host[k % R], ref_host[k % R], H, W, loss_host[k % Se], gmvp_host[k % Se])

    n_up = min(e2e_steps, 400)
    ms_up = e2e_time(begin_u8, n_up)
    e2e["with_mask_upload_every_step"] = {"value": B * n_up * world / (ms_up * 1e-3), "unit": "frames/s",
                                          "h2d_bytes_per_step": int(B * H * W + B * L * 64), "steps": n_up,
                                          "api": "ehb_solver_step_begin_u8 / _end (u8 masks + mvp in, every step)"}

    if rank == 0:
        cpu, parity, proxy, nvd = None, None, None, None
        if world == 1 and not args.no_cpu:
            from oracle import oracle
            oracle.build()
            # CPU leg: the oracle on the very views that were timed (every ring slot), 5 passes over the ring; its
            # results are the checker of the GPU results above
            fps, cms, threads, last = cpu_reference_run(wl, sets, 5 * R, R)
            cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": "the %d ring slots (%d views each) x 5 passes (+1 warm-up pass), OpenMP over views" % (R, B)}
            ok, worst_g, worst_l = True, 0.0, 0.0
            for s in range(R):
                gm, gl, gg, _ = gpu_results[s]
                w = last[s]
                ok &= bool(np.array_equal(gm, w["masks"]))
                worst_l = max(worst_l, rel_err(gl, w["loss_per_view"]))
                worst_g = max(worst_g, rel_err(gg, w["g_mvp"]))
            ok &= worst_l < 1e-12 and worst_g < 1e-6
            if slot_path_ok is not None:
                ok &= slot_path_ok
            parity = {"ok": bool(ok), "slots": R, "masks": "bit-exact" if ok else "MISMATCH", "loss_rel_err": worst_l,
                      "g_mvp_rel_err": worst_g, "checker": "oracle.render_views on the timed inputs (outside the timed region)",
                      "slot_path_equals_single_stream_path": slot_path_ok}
            if not args.no_proxy:
                try:
                    proxy = time_proxy(wl, sets, ids, dev)
                except Exception as e:   # noqa: BLE001
                    proxy = "failed: %s" % str(e)[:200]
                nvd = time_nvdiffrast(wl, sets, dev)
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "views_per_step_per_gpu": B, "H": H, "W": W, "links": L,
                           "triangles": F, "vertices": V, "onscreen_frac": onscreen, "coverage": coverage,
                           "step": "fused render + loss + backward to d loss/d mvp, pose chain to d loss/d dof, " +
                                   ("all-reduce of 7 floats, " if world > 1 else "") + "Adam",
                           "l2": "ring of %d view-sets (%.0f MB of masks+refs) > 126 MB L2" %
                                 (R, R * B * H * W * 8 / 1e6),
                           "pipelines": pipes_used,
                           "launch": ("%d steps in flight on the context's slots (ehb_step_begin), " % S +
                                      ("CUDA graph replay, %d steps per graph" % G if slot_graph is not None else "eager"))
                                     if inflight else ("CUDA graph replay, one graph per ring slot" if graphs is not None else "eager"),
                           "steps_in_flight": S if inflight else 1,
                           "scenes": "identical on every rank", "collective": collective,
                           "timed_regions": len(regions), "timed_region_ms": ms},
                "clocks": clocks, "e2e": e2e, "serial": serial, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "parity_checked": bool(parity and parity["ok"]), "parity": parity,
                "proxy": proxy, "nvdiffrast": nvd, "pytorch3d": probe_pytorch3d(),
                "need_clip_triangles": int(nclip)}
        out.emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS) + ["explore"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity / proxy legs")
    ap.add_argument("--no-proxy", action="store_true", help="skip the unfused-loop proxy and the nvdiffrast probe")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"], help="N>1: how the 7 floats are all-reduced")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "explore":
        run_explore(args)
    else:
        run_b200(args)


def run_explore(args):
    """BASELINE.json configs[3]: space-exploration scoring, 256 candidate joint configurations x 4 camera poses, xArm7 base +
    links 1-7 (41,096 triangles), 1920x1080, forward-only binary render + variance score.  Candidates are block-partitioned
    over the ranks (no data-path collective), one all-gather of the scores ends a step.  value = renders/s of the job."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = JsonStdout(world > 1)
    import pathlib
    import tempfile
    import torch
    import torch.distributed as dist
    from easyhec_b200._lib import Context
    from easyhec_b200.explore import score_candidates, shard_candidates
    from easyhec_b200.scenes import FRANKA_K, load_xarm7, make_scene, perturb_pose, scaled_K
    from util import xarm_urdf
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ex = EXPLORE
    Q, C, H, W = ex["Q"], ex["C"], ex["H"], ex["W"]
    fx = load_xarm7()
    kin = xarm_urdf(pathlib.Path(tempfile.mkdtemp()), fx)
    lim = fx["joint_limits"]
    q = np.random.RandomState(0).uniform(np.maximum(lim[:, 0], -np.pi) * 0.6, np.minimum(lim[:, 1], np.pi) * 0.6, size=(Q, len(lim)))
    sc = make_scene(1, H, W, links="xarm7_all", seed=0, K_base=FRANKA_K)
    cams = np.stack([perturb_pose(sc["Tc_c2b"], np.random.RandomState(1 + c), 0.05, 5.0) for c in range(C)])
    K = scaled_K(H, W, FRANKA_K)
    ctx = Context(dev)
    ids = [ctx.register_mesh(m.vertices, m.faces) for m in fx["meshes"]]
    robot = ctx.register_robot(kin)
    links = list(range(8))
    V = sum(len(m.vertices) for m in fx["meshes"]); F = sum(len(m.faces) for m in fx["meshes"])
    q_dev = torch.from_numpy(q).to(dev)

    def step(_k):
        return score_candidates(ctx, ids, kin, links, q_dev, cams, K, H, W, robot=robot)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(max(args.warmup, 3)):
        scores = step(k)
    flags, nclip = ctx.status()
    assert flags & 1 == 0
    sampler = ClockSampler(physical_gpu_index(local))
    barrier()
    sampler.start()
    l0 = ctx.launch_count()
    steps = min(args.steps, 50)
    def agree(done):               # every rank runs the same number of timed regions
        if world == 1:
            return done
        t = torch.tensor([1.0 if done else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() >= 1.0)

    ms, regions = timed_regions(lambda n, k0: [step(k0 + k) for k in range(n)], steps, barrier, torch, agree=agree)
    clocks = sampler.stop()
    launches = (ctx.launch_count() - l0) // max(len(regions), 1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    renders = Q * C * steps
    value = renders / (ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = algorithmic_bytes_per_render(H, W, V, F) * Q * C        # per step of the whole job
    if rank == 0:
        parity = None
        if not args.no_cpu and world == 1:
            from oracle import oracle
            packed = oracle.pack_links(fx["meshes"])
            n = 8                                                   # a bounded sample of the candidates on the host cores
            mvp = ctx.explore_fk_mvp(robot, q_dev[:n].contiguous(), cams, K, H, W, links).cpu().numpy()
            t0 = time.perf_counter()
            masks = oracle.union_binary(packed, mvp.reshape(n * C, len(links), 4, 4), H, W)
            want = oracle.variance_scores(masks.reshape(n, C, H, W))
            dt = time.perf_counter() - t0
            parity = {"ok": bool(np.allclose(scores[:n].cpu().numpy(), want, rtol=1e-14, atol=0)), "candidates": n,
                      "cpu_renders_per_s": n * C / dt, "cores": oracle.num_threads()}
        line = {"metric": "space-exploration scoring renders/sec @1920x1080 xArm7 (256 qpos x 4 cameras)", "value": value,
                "unit": "renders/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": ex["name"], "Q": Q, "C": C, "H": H, "W": W, "triangles": F, "candidates_per_rank":
                           shard_candidates(Q, 0, world)[1], "collective": "one all-gather of the scores per step" if world > 1 else "none",
                           "step": "device FK->MVP, packed binary render, variance from the depth planes" +
                                   (", all-gather of %d scores" % Q if world > 1 else ""),
                           "timed_regions": len(regions), "timed_region_ms": ms},
                "candidates_per_s": Q * steps / (ms * 1e-3), "clocks": clocks, "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": alg / (ms / steps * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                             "frac": alg / (ms / steps * 1e-3) / 1e9 / (peak * world), "algorithmic_bytes_per_step": alg,
                             "note": "whole step against the aggregate peak of the GPUs used"},
                "parity_checked": bool(parity and parity["ok"]), "parity": parity, "best_candidate": int(scores.argmax().item())}
        out.emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
