#!/usr/bin/env python
"""bench.py — gradient-carrying path samples/sec of the EPSM manifold path on N B200s (BASELINE.json metric).

A "step" is one gradient pass (render_backward-equivalent: ray generation, sample_path replay, manifold solve,
gradient scatter; epsm.py:84-306) over the W·H·spp wavefront of the workload, followed — for N > 1 — by the one
NCCL allreduce of the mesh-gradient buffer.  Work is sharded across ranks by sample index.  Default `--scaling strong`:
the workload's wavefront (BASELINE configs[2]: 1024²×32 spp, a fixed job) is split, each rank tracing spp/N samples of every
pixel; the weak-scaling number (every rank traces spp samples per pixel, i.e. an (N·spp)-sample wavefront — BASELINE's
resolution/spp sweep) is measured in the same run and printed under the key `weak`.  `--scaling weak` makes that the headline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload room|glass|cbox|smpl]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N … bench.py --gpus N …

Prints ONE JSON line (rank 0).  `value` = device-resident inputs, CUDA-event time on the library's stream, max over
ranks; `e2e` = the same pass through the C ABI with HOST buffers (pinned grad_in H2D, gradient buffers D2H inside
the timed region); `roofline` = algorithmic traversal bytes of the gradient kernel ÷ its measured duration ÷ the
measured HBM copy bandwidth; `cpu_baseline` = the CPU oracle (a port of the reference's path; the reference itself
needs cmake + a CUDA-only torch path and cannot be built from a few files) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (builder kwargs, variant, max_depth, description)
    "room": dict(variant="manifold", max_depth=8, res=1024, spp=32,
                 desc="procedural 1M-triangle indoor scene, manifold(_hybrid phase 1), 1024x1024, 32 spp (BASELINE configs[2])"),
    "room8m": dict(variant="manifold", max_depth=8, res=1024, spp=32, tris=8_000_000,
                   desc="the procedural indoor scene at 8M triangles (~0.8 GB of BVH nodes + triangle slots: does NOT fit the 126 MB L2), manifold, 1024x1024, 32 spp"),
    "glass": dict(variant="manifold_caustic", max_depth=6, res=512, spp=64,
                  desc="glass icosphere (20480 tris) over diffuse plane + area light, manifold_caustic, 512x512, 64 spp (configs[1])"),
    "cbox": dict(variant="manifold", max_depth=6, res=128, spp=16, desc="mi.cornell_box() 128x128, 16 spp, manifold (configs[0])"),
    "smpl": dict(variant="manifold_caustic", max_depth=3, res=512, spp=64,
                 desc="6890-vertex articulated mesh stage, manifold_caustic, 512x512, 64 spp (configs[3])"),
}


def build_scene(name, res, spp, tris):
    from epsm_mitsuba3_b200 import scenes as S
    if name in ("room", "room8m"):
        sd = S.indoor_room(tris, res, spp, res)
    elif name == "glass":
        sd = S.glass_sphere_caustic(res, spp, res)
    elif name == "cbox":
        sd = S.cornell_box(res, spp, res, box_bsdf=S.Bsdf.conductor())
    elif name == "smpl":
        sd = S.articulated_stage(res, spp, res)
    else:
        raise SystemExit(f"unknown workload {name}")
    return sd


def synthetic_grad_in(res, seed=7):
    rng = np.random.default_rng(seed)                      # matcher stand-in of SURVEY.md §8d
    g = np.zeros((res, res, 5), np.float32)
    g[..., :3] = rng.normal(0, 0.1, (res, res, 3)); g[..., 3:] = rng.normal(0, 0.02, (res, res, 2))
    return g


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            v = d.get("hbm_gbs") or d.get("hbm_gb_s")
            if v:
                return float(v), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def iteration_leg(ep, scene, sd, integ, wl, sensor, spp, g_dev, buf, n=5, warm=2):
    import torch
    params = ep.traverse(scene)
    keys = [k for k, i in params._mesh_of.items() if sd.meshes[i].grad] or list(params._mesh_of)
    base = {k: params[k].clone() for k in keys}
    out = {"unit": "ms per iteration (wall, median)", "iterations": n, "meshes_updated": len(keys),
           "vertices_updated": int(sum(base[k].shape[0] for k in keys)),
           "what": "params.update() of every differentiable mesh (device tensors) + vertex normals + BVH refresh + primal render "
                   "(sensor 1, same res/spp) + Sinkhorn matcher at match_res (native kernels in place of third-party geomloss; its "
                   "gradient, tiled to the film, is the grad_in of the pass) + gradient pass + Adam-style vertex step"}
    scene.set_sample_range(0, 0)
    # the matcher (Matcher.match_Sinkhorn, EPSM/utils/matcher.py:51-63; optim.py:130-136): rendered image → match_res² →
    # 5-channel gradient tiled back to the film size.  Target = the render of the unperturbed scene.
    import torch.nn.functional as F
    match_res = 128
    matcher = ep.Matcher(match_res, scene.device)
    res = sd.sensors[1].width

    def low(img):
        return F.adaptive_avg_pool2d(img[..., :3].permute(2, 0, 1)[None], match_res)[0].permute(1, 2, 0).reshape(-1, 3)
    target_low = low(integ.render(scene, sensor=1, seed=12345, spp=spp)).clone()
    out["matcher"] = {"match_res": match_res, "points": match_res * match_res}
    for mode in ("refit", "rebuild"):
        scene.accel_update = mode
        ts, parts = [], []
        for it in range(warm + n):
            buf.zero_(); torch.cuda.synchronize(); scene.synchronize()
            t0 = time.perf_counter()
            for k in keys:
                params[k] = base[k] + 1e-3 * float(np.sin(it + 1.0))
            params.update(); scene.synchronize()
            t1 = time.perf_counter()
            img = integ.render(scene, sensor=1, seed=it, spp=spp); scene.synchronize(); torch.cuda.synchronize()
            t2 = time.perf_counter()
            g5 = matcher.match_Sinkhorn(low(img), target_low).reshape(match_res, match_res, 5)
            g_it = g5.repeat_interleave(res // match_res, 0).repeat_interleave(res // match_res, 1).contiguous() if res % match_res == 0 else g_dev
            torch.cuda.synchronize()
            tm = time.perf_counter()
            scene.render_backward_raw(wl["variant"], sensor, ep.sample_tea_32(it, 1)[0], spp, integ.max_depth, integ.rr_depth, g_it, buf[0], buf[1])
            scene.synchronize()
            t3 = time.perf_counter()
            voff = scene.voff
            for k in keys:
                i = params._mesh_of[k]
                base[k].add_(buf[0][int(voff[i]):int(voff[i + 1])].reshape(base[k].shape), alpha=0.0)   # optimiser step stand-in (lr 0: the scene stays put)
            torch.cuda.synchronize()
            t4 = time.perf_counter()
            if it >= warm:
                ts.append((t4 - t0) * 1e3); parts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (tm - t2) * 1e3, (t3 - tm) * 1e3, (t4 - t3) * 1e3))
        pm = np.median(np.array(parts), axis=0)
        out[mode] = {"ms": float(np.median(ts)), "ms_without_matcher": float(np.median(ts) - pm[2]), "update_ms": float(pm[0]), "primal_ms": float(pm[1]),
                     "matcher_ms": float(pm[2]), "backward_ms": float(pm[3]), "step_ms": float(pm[4])}
    out["matcher"]["eps_steps"] = matcher.n_steps
    for k in keys:                                            # leave the scene as it was built
        params[k] = base[k]
    scene.accel_update = "rebuild"; params.update(); scene.accel_update = "auto"
    return out


def phase2_leg(ep, scene, sd, wl, spp, n=2):
    """ms per optimisation iteration once the hybrid schedule has switched (optim.py:113-141): vertex update + BVH refresh, primal render of
    sensor 0 (sample_border) with prb_reparam's estimator, L2 loss against a target, reverse pass (k_prb_backward: primal film, film adjoint,
    16 auxiliary rays per reparameterised ray), vertex step."""
    import torch
    integ2 = ep.load_dict({"type": "prb_reparam", "max_depth": wl["max_depth"]})
    params = ep.traverse(scene)
    keys = [k for k, i in params._mesh_of.items() if sd.meshes[i].grad] or list(params._mesh_of)
    base = {k: params[k].clone() for k in keys}
    target = integ2.render(scene, sensor=0, seed=999, spp=spp).clone()
    ts, parts = [], []
    for it in range(1 + n):
        torch.cuda.synchronize(); scene.synchronize()
        t0 = time.perf_counter()
        for k in keys:
            params[k] = base[k] + 1e-3 * float(np.sin(it + 1.0))
        params.update(); scene.synchronize()
        t1 = time.perf_counter()
        img = integ2.render(scene, sensor=0, seed=it, spp=spp); scene.synchronize()
        t2 = time.perf_counter()
        g = (2.0 * (img - target) / img.numel()).contiguous()                  # d/d img of sum(sqr(img − ref)) / len(img)  (optim.py:138)
        gp, gn = integ2.vertex_gradients(scene, g, 0, ep.sample_tea_32(it, 1)[0], spp)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if it >= 1:
            ts.append((t3 - t0) * 1e3); parts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
    for k in keys:
        params[k] = base[k]
    params.update()
    pm = np.median(np.array(parts), axis=0)
    return {"unit": "ms per iteration (wall, median)", "iterations": n, "integrator": "prb_reparam", "sensor": 0, "spp": spp, "reparam_rays": integ2.reparam_rays,
            "ms": float(np.median(ts)), "update_ms": float(pm[0]), "primal_ms": float(pm[1]), "backward_ms": float(pm[2]),
            "grad_abs_sum": float(gp.abs().sum())}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpus=(0,), enabled=True):
        """One nvidia-smi process for all the job's GPUs, started by rank 0 only (`enabled`): eight of them polling the driver
        every 200 ms from eight ranks showed up as multi-millisecond outliers in the 8-GPU strong-scaling steps."""
        super().__init__(daemon=True); self.gpus = list(gpus); self.rows = []; self.stop_evt = threading.Event(); self.proc = None
        self.t_begin = None; self.enabled = enabled

    def wait_first(self, timeout):
        if not self.enabled:
            return
        t0 = time.time()
        while not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def begin(self):
        self.t_begin = time.time()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        if not self.enabled:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--id=" + ",".join(str(g) for g in self.gpus), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")] + [time.time()])
                if self.stop_evt.is_set():
                    break
        except Exception:
            pass

    def finish(self):
        t_end = time.time()
        self.stop_evt.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        rows = [r for r in self.rows if self.t_begin is None or self.t_begin <= r[-1] <= t_end + 0.05]
        if not rows:
            # timed region shorter than the 200 ms sampling period: the last row before its end (the GPU has been under
            # the same load since the warm-up steps)
            rows = [r for r in self.rows if r[-1] <= t_end + 0.05][-1:]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "gpus": self.gpus}


def cpu_port_throughput(sd, wl, grad_in, target_seconds=12.0, prune=True):
    """The CPU oracle (port of the reference's path) on a bounded sample of the workload, all host threads.
    prune=True: the replay stops where each path's gradient chain ends — the same work the GPU pass does (its default);
    prune=False: the reference's full min(max_depth, 6)-iteration replay (compare with the GPU line's `full_replay`)."""
    import ctypes as C
    import torch
    from oracle import oracle
    from oracle.calc_grad_ref import calc_grad, scatter_gradients
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    o = oracle.OracleScene(sd, accel=True)
    oracle.lib().orc_set_prune.argtypes = [C.c_int]
    oracle.lib().orc_set_prune(({"manifold": 0, "manifold_caustic": 1}[wl["variant"]]) if prune else -1)
    sensor, spp, md = 1, wl["spp"], wl["max_depth"]
    total = sd.sensors[sensor].width * sd.sensors[sensor].height * spp
    nl = min(min(md, 6), 5)

    def run(n):
        t0 = time.perf_counter()
        # a contiguous block of whole pixels from the middle of the image
        lb = (total // 2 // spp) * spp; le = min(total, lb + n)
        paths = o.path_log(sensor, 1, spp, md, 5, grad_in, lb, le)
        pg, lg, dg = calc_grad(paths, wl["variant"], nl)
        scatter_gradients(paths, wl["variant"], nl, pg, lg, dg, sd.n_vertices, md)
        return le - lb, time.perf_counter() - t0

    try:
        n, dt = run(4096)
        rate = n / dt
        n2 = int(min(total // 4, max(4096, rate * target_seconds)))
        n, dt = run(n2)
    finally:
        oracle.lib().orc_set_prune(-1)
    return {"value": n / dt, "unit": "gradient path samples/s", "cores": cores, "kind": "port",
            "replay": "gradient-live prefix (same work as the GPU pass)" if prune else "full (the reference's min(max_depth,6) iterations per lane)",
            "sample": f"{n} consecutive lanes (whole pixels, image centre) of the same wavefront: oracle path_log (OpenMP, median-split BVH) "
                      f"+ torch calc_grad + NumPy scatter; {dt:.1f} s"}


def reference_binary_throughput(sd, wl, res_sample=256, spp_sample=None, timeout=600):
    """BASELINE.md §2 B-trav: the UNMODIFIED reference (baseline/_ref/{kd,embree}/mitsuba, scalar_rgb, built by baseline/build_ref.sh)
    rendering the same scene with its `path` integrator on all host cores — the traversal-dominated primal half of the path
    (the scalar build has no AD, so the manifold pass itself cannot run in it).  Bounded sample: res_sample² pixels at the
    workload's spp.  Returns {variant: {...}}; entries say why when a binary is missing."""
    import re
    import shutil
    import tempfile
    from epsm_mitsuba3_b200 import export
    cores = os.cpu_count() or 1
    spp = spp_sample or wl["spp"]
    out = {}
    tmp = tempfile.mkdtemp(prefix="epsm_refscene_")
    try:
        xml = export.export_scene(sd, tmp, integrator={"type": "path", "max_depth": int(wl["max_depth"])}, sensors=[1], spp=spp, resolution=res_sample)
        for variant, accel in (("kd", "SAH kd-tree (kdtree.h:2188-2290)"), ("embree", "Embree 3 (scene_embree.inl:202-426)")):
            d = os.path.join(ROOT, "baseline", "_ref", variant)
            exe = os.path.join(d, "mitsuba")
            if not os.path.exists(exe):
                out[variant] = {"unavailable": f"{exe} not built (baseline/build_ref.sh{' --embree' if variant == 'embree' else ''})"}
                continue
            env = dict(os.environ, LD_LIBRARY_PATH=d + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
            try:
                r = subprocess.run([exe, "-m", "scalar_rgb", "-t", str(cores), "-o", os.path.join(tmp, f"{variant}.exr"), xml], env=env,
                                   capture_output=True, text=True, timeout=timeout)
            except subprocess.TimeoutExpired:
                out[variant] = {"unavailable": f"timed out after {timeout} s"}; continue
            log = r.stdout + r.stderr

            def secs(pat):
                m = re.search(pat, log)
                if not m:
                    return None
                v, u = float(m.group(1)), m.group(2)
                return v * {"ms": 1e-3, "s": 1.0, "m": 60.0}[u]
            t_render = secs(r"Rendering finished\. \(took ([0-9.]+)(ms|s|m)")
            t_build = secs(r"(?:Finished\. \([^)]*took|Embree ready\. \(took) ([0-9.]+)(ms|s|m)")
            if r.returncode != 0 or not t_render:
                out[variant] = {"unavailable": f"exit code {r.returncode}: {log[-300:]}"}; continue
            n = res_sample * res_sample * spp
            out[variant] = {"value": n / t_render, "unit": "primal path samples/s", "cores": cores, "accel": accel, "accel_build_s": t_build,
                            "render_s": t_render, "sample": f"`mitsuba -m scalar_rgb -t {cores}` path integrator, max_depth {wl['max_depth']}, {res_sample}x{res_sample}, {spp} spp of the exported scene"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="room", choices=sorted(WORKLOADS))
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--tris", type=int, default=1_000_000)
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): the workload's spp-sample wavefront — BASELINE configs[2] is a fixed 1024x1024x32 job — is split over the ranks; "
                         "weak: per-GPU work fixed (N ranks trace N*spp samples per pixel, BASELINE's resolution/spp sweep)")
    ap.add_argument("--no-weak-leg", action="store_true", help="N > 1, strong scaling: skip the weak-scaling measurement printed beside the line")
    ap.add_argument("--phase2-leg", action="store_true", help="also time one optimisation iteration of hybrid phase 2 (prb_reparam on sensor 0, L2 loss)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-iteration-leg", action="store_true", help="skip the ms-per-optimisation-iteration measurement (BASELINE.json's second metric)")
    ap.add_argument("--full-replay", action="store_true", help="time the unpruned replay (every lane runs min(max_depth,6) iterations, as the reference does)")
    ap.add_argument("--no-full-replay-leg", action="store_true", help="skip the extra full-replay measurement reported beside the default")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = dict(WORKLOADS[args.workload])
    if "tris" in wl and args.tris == 1_000_000:
        args.tris = wl["tris"]
    if args.res: wl["res"] = args.res
    if args.spp: wl["spp"] = args.spp
    res, spp = wl["res"], wl["spp"]
    spp_total = spp * world if args.scaling == "weak" else spp       # samples per pixel of the whole job
    config = {"workload": wl["desc"], "scene": args.workload, "integrator": wl["variant"], "max_depth": wl["max_depth"],
              "resolution": [res, res], "spp": spp_total, "spp_per_gpu": spp_total // world, "paths_per_step": res * res * spp_total,
              "sharding": f"sample-index shards over {world} rank(s) ({args.scaling} scaling: {spp_total // world} of the {spp_total} samples of every pixel per rank) "
                          f"+ 1 NCCL allreduce of the vertex-gradient buffer" if world > 1 else "single GPU",
              "l2": "256 MiB buffer written between timed steps (L2 flush); BVH+triangles of the 1M-tri scene are ~100 MB",
              "replay": "full (every lane replays min(max_depth,6) iterations)" if args.full_replay else
                        "gradient-live prefix (each path stops where calc_grad masks everything behind it; gradients identical to the full replay)"}

    if args.impl == "reference":
        if rank != 0:
            return
        sd = build_scene(args.workload, res, spp, args.tris)
        g = synthetic_grad_in(res)
        vals = []
        for i in range(args.warmup + args.steps):
            r = cpu_port_throughput(sd, wl, g, target_seconds=8.0, prune=not args.full_replay)
            if i >= args.warmup:
                vals.append(r)
        v = float(np.mean([r["value"] for r in vals])) if vals else 0.0
        r = vals[-1] if vals else {"cores": os.cpu_count(), "kind": "port", "sample": "none"}
        full = cpu_port_throughput(sd, wl, g, target_seconds=8.0, prune=False) if not args.full_replay else None
        try:
            refbin = reference_binary_throughput(sd, wl)
        except Exception as e:
            refbin = {"unavailable": f"{type(e).__name__}: {e}"}
        out = {"impl": "reference", "metric": "gradient path samples/sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": (res * res * spp / v * 1e3) if v else None,
               "ms_per_step_note": "one whole pass of the workload at the sampled rate (each step times a bounded sample, see cpu_baseline.sample)",
               "higher_is_better": True, "scaling": args.scaling,
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
               "cpu_baseline": {"value": v, "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"], "replay": r.get("replay")},
               "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "full_replay": ({"value": full["value"], "unit": "samples/s", "sample": full["sample"],
                                "note": "the reference's own schedule (no pruning); compare with the GPU line's full_replay key"} if full else None),
               "reference_binary": refbin,
               "note": "kind = port: the reference's manifold path is Python + Dr.Jit AD variants + CUDA-only torch (SURVEY F4, §8c); the scalar_rgb build "
                       "of the reference that exists here (baseline/_ref) has no AD and cannot run it.  `value` times the CPU oracle — a line-by-line "
                       "port of the same path — on all host cores, on the same pruned schedule the GPU arm runs; `reference_binary` is the unmodified "
                       "reference binary's primal path tracer (kd-tree and Embree) on the exported scene: BASELINE.md §2's B-trav"}
        print(json.dumps(out)); return

    import torch
    import torch.distributed as dist
    import epsm_mitsuba3_b200 as ep
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sd = build_scene(args.workload, res, spp, args.tris)
    scene = ep.Scene(sd, device=local)
    scene.set_full_replay(args.full_replay)
    stats = scene.stats()
    integ = ep.load_dict({"type": wl["variant"], "max_depth": wl["max_depth"]})
    sensor = 1                                                # full-resolution sensor; benchmark semantics N_g = W·H·spp
    g_host = torch.from_numpy(synthetic_grad_in(res)).pin_memory()
    g_dev = g_host.cuda()
    spp_job = spp_total                                       # lane ids are pixel·spp_job + s for every rank
    s0, s1 = ep.sample_shard(spp_job, rank, world)
    buf = torch.zeros((2, scene.n_vertices, 3), dtype=torch.float32, device=scene.device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=scene.device)
    torch.cuda.synchronize()
    import ctypes as C
    cu_stream = scene.stream()

    def step(k, timed):
        """One gradient pass over this rank's shard (+ allreduce).  Returns device ms of the pass."""
        buf.zero_(); flush.fill_(k & 0xff)                   # L2 flush, outside the event pair
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(cu_stream)
        scene.set_sample_range(s0, s1)
        scene.render_backward_raw(wl["variant"], sensor, 100 + k, spp_job, integ.max_depth, integ.rr_depth, g_dev, buf[0], buf[1])
        ev1.record(cu_stream)
        scene.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2.record(); dist.all_reduce(buf); e3.record(); torch.cuda.synchronize()
            ms += e2.elapsed_time(e3)
        return ms

    clk = ClockSampler(list(range(world)) if world > 1 else [local], enabled=(rank == 0)); clk.start()                  # nvidia-smi takes a while to start on an 8-GPU box: launch it before the warm-up
    if world > 1:                                            # communicator set-up (NCCL channels, NVLS buffers) is not a step
        for _ in range(3):
            dist.all_reduce(buf)
        torch.cuda.synchronize()
    for k in range(args.warmup):
        step(k, False)
    clk.wait_first(2.0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clk.begin()
    launches0 = scene.launch_count()
    t_wall0 = time.perf_counter()
    ms_steps = [step(args.warmup + k, True) for k in range(args.steps)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = scene.launch_count() - launches0
    clocks = clk.finish()
    ms_total = float(np.sum(ms_steps))
    if world > 1:
        t = torch.tensor([ms_total], device=scene.device); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_total = float(t.item())
        l = torch.tensor([launches], device=scene.device); dist.all_reduce(l); launches = int(l.item())
    ms_per_step = ms_total / args.steps
    total_paths = res * res * spp_job
    value = total_paths / (ms_per_step * 1e-3)

    # ---- the unpruned replay beside the default, for transparency (same scene, same seeds)
    full_leg = None
    if not args.full_replay and not args.no_full_replay_leg:
        scene.set_full_replay(True)
        step(0, False)
        fm = [step(args.warmup + k, True) for k in range(2)]
        scene.set_full_replay(False)
        fms = float(np.mean(fm))
        if world > 1:
            t = torch.tensor([fms], device=scene.device); dist.all_reduce(t, op=dist.ReduceOp.MAX); fms = float(t.item())
        full_leg = {"value": total_paths / (fms * 1e-3), "unit": "samples/s", "ms_per_step": fms,
                    "note": "epsm_set_full_replay(1): no gradient-liveness pruning, 1 warm-up + 2 timed steps"}

    # ---- weak scaling beside the strong headline (N > 1): every rank traces the workload's spp samples of every pixel
    weak_leg = None
    if world > 1 and args.scaling == "strong" and not args.no_weak_leg:
        spp_w = spp * world
        w0, w1 = ep.sample_shard(spp_w, rank, world)

        def wstep(k):
            buf.zero_(); flush.fill_(k & 0xff); torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(cu_stream)
            scene.set_sample_range(w0, w1)
            scene.render_backward_raw(wl["variant"], sensor, 100 + k, spp_w, integ.max_depth, integ.rr_depth, g_dev, buf[0], buf[1])
            ev1.record(cu_stream); scene.synchronize()
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2.record(); dist.all_reduce(buf); e3.record(); torch.cuda.synchronize()
            return ev0.elapsed_time(ev1) + e2.elapsed_time(e3)
        wstep(0); dist.barrier()
        wm = float(np.mean([wstep(1 + k) for k in range(3)]))
        t = torch.tensor([wm], device=scene.device); dist.all_reduce(t, op=dist.ReduceOp.MAX); wm = float(t.item())
        weak_leg = {"value": res * res * spp_w / (wm * 1e-3), "unit": "samples/s", "ms_per_step": wm, "spp": spp_w, "spp_per_gpu": spp,
                    "note": "weak scaling: per-GPU work fixed (the N-rank job is an N*spp-sample wavefront); 1 warm-up + 3 timed steps, max over ranks"}
        scene.set_sample_range(s0, s1)

    # ---- e2e: the same pass through the C ABI with host buffers (pinned H2D of grad_in, D2H of the gradient buffers)
    e2e = None
    if not args.no_e2e:
        gp_h = torch.zeros((scene.n_vertices, 3), dtype=torch.float32).pin_memory(); gn_h = torch.zeros_like(gp_h).pin_memory()
        scene.set_sample_range(s0, s1)
        if world > 1:
            g_chunk = (g_host.numel() + world - 1) // world
            g_host_flat = torch.zeros(g_chunk * world, dtype=torch.float32).pin_memory(); g_host_flat[:g_host.numel()] = g_host.reshape(-1)
            g_slice = torch.empty(g_chunk, dtype=torch.float32, device=scene.device); g_full = torch.empty(g_chunk * world, dtype=torch.float32, device=scene.device)
        ts = []
        for k in range(2 + args.steps):
            gp_h.zero_(); gn_h.zero_(); flush.fill_(k & 0xff); torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            if world == 1:
                scene.render_backward_raw(wl["variant"], sensor, 200 + k, spp_job, integ.max_depth, integ.rr_depth, g_host, gp_h, gn_h)   # host pointers through the C ABI: H2D + pass + D2H
            else:
                # N ranks: every rank uploads 1/N of the pinned grad_in over its own PCIe link and the slices are all-gathered
                # over NVLink (the whole image is needed by every rank: shards are by sample index), device pass, ONE
                # allreduce of the device buffer, ONE D2H of the reduced gradients
                g_slice.copy_(g_host_flat[rank * g_chunk:(rank + 1) * g_chunk], non_blocking=True)
                dist.all_gather_into_tensor(g_full, g_slice)
                g_d = g_full[:g_host.numel()].view(g_host.shape); buf.zero_(); torch.cuda.current_stream().synchronize()
                scene.render_backward_raw(wl["variant"], sensor, 200 + k, spp_job, integ.max_depth, integ.rr_depth, g_d, buf[0], buf[1])
                scene.synchronize()
                dist.all_reduce(buf)
                gp_h.copy_(buf[0], non_blocking=True); gn_h.copy_(buf[1], non_blocking=True); torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if k >= 2:
                ts.append(dt)
        te = float(np.mean(ts))
        if world > 1:
            t = torch.tensor([te], device=scene.device); dist.all_reduce(t, op=dist.ReduceOp.MAX); te = float(t.item())
        h2d = g_host.numel() * 4 + (2 * gp_h.numel() * 4 if world == 1 else 0); d2h = 2 * gp_h.numel() * 4     # C ABI accumulates: output buffers go up too
        e2e = {"value": total_paths / te, "unit": "samples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": te * 1e3}
    scene.set_sample_range(s0, s1)

    # ---- roofline of the dominant kernel (k_wf_trace, the persistent BVH traversal): algorithmic bytes from the
    #      instrumented variant of the same kernel (identical control flow), duration from CUDA events that bracket
    #      every traversal launch of one un-instrumented pass on the library's stream.
    cnt = np.zeros(4, np.uint64)
    buf.zero_(); torch.cuda.synchronize()
    scene.render_backward_raw(wl["variant"], sensor, 100 + args.warmup, spp_job, integ.max_depth, integ.rr_depth, g_dev, buf[0], buf[1], counters=cnt)
    scene.synchronize()
    nodes, tris, rays, lanes = (int(x) for x in cnt)
    alg_bytes = nodes * stats["node_bytes"] + tris * stats["tri_bytes"] + rays * 48
    scene.set_trace_profiling(True)
    scene.trace_profile()
    buf.zero_(); flush.fill_(1); torch.cuda.synchronize()
    scene.render_backward_raw(wl["variant"], sensor, 100 + args.warmup, spp_job, integ.max_depth, integ.rr_depth, g_dev, buf[0], buf[1])
    trace_ms, trace_launches = scene.trace_profile()
    scene.set_trace_profiling(False)
    peak, peak_src = peaks()
    roof = None
    if rank == 0:
        k_ms = float(np.mean(ms_steps))
        ach = alg_bytes / (trace_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.workload)
            except Exception:
                traffic = None
        avg_launch_s = trace_ms * 1e-3 / max(trace_launches, 1)
        scene_bytes = int(stats["n_nodes"]) * int(stats["node_bytes"]) + int(stats["n_triangles"]) * 64
        roof = {"bound": "hbm", "kernel": "k_wf_trace (closest + any-hit launches of one gradient pass)", "achieved": ach, "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per traversal launch from the committed ncu "
                                  "--set full capture of this workload (a constant of the capture, NOT re-measured in this run; null for workloads never captured)",
                "dram_GBps_from_capture": (traffic / avg_launch_s / 1e9) if traffic else None,
                "dram_frac_of_peak_from_capture": (traffic / avg_launch_s / 1e9 / peak) if traffic else None,
                "scene_bytes_nodes_plus_triangle_slots": scene_bytes, "l2_bytes": 126 * 1024 * 1024,
                "reading": "`frac` is the contract's figure: ALGORITHMIC bytes (cache hits not subtracted) over the measured copy bandwidth.  When the BVH "
                           "and triangle slots fit the 126 MB L2 (scene_bytes <= l2_bytes) the kernel is NOT HBM-bound and frac can exceed 1: DRAM moves "
                           "only the ray/hit streams (dram_frac_of_peak_from_capture); the binding units are then the L1 data pipe and the ALU (FMNMX) "
                           "pipe — ncu figures in profiles/.  `--workload room8m` (8 M triangles, ~0.8 GB of nodes + slots) is the case that does not fit L2",
                "launches_per_step": trace_launches, "kernel_ms_per_step": trace_ms, "share_of_step": trace_ms / k_ms,
                "algorithmic_bytes_per_step": alg_bytes, "algorithmic_bytes_per_launch": alg_bytes / max(trace_launches, 1),
                "avg_launch_ms": trace_ms / max(trace_launches, 1),
                "whole_step_GBps": alg_bytes / (k_ms * 1e-3) / 1e9,
                "per_path": {"rays": rays / max(lanes, 1), "nodes": nodes / max(lanes, 1), "tris": tris / max(lanes, 1),
                             "bytes": alg_bytes / max(lanes, 1)},
                "definition": "nodes_fetched*128 + tris_tested*48 + rays*(32+16) bytes (SURVEY 8d), counted by the instrumented kernel on the "
                              "same seed, divided by the summed CUDA-event duration of the traversal launches; cache hits are not subtracted"}

    # ---- ms per optim.py iteration (BASELINE.json second metric; EPSM/optim.py:110-157, matcher included and reported apart):
    #      vertex update of every differentiable mesh → normals + BVH refresh → primal render (sensor 1) → gradient pass →
    #      Adam-style step on the vertex tensors.  Wall clock, median, once per BVH policy.
    iteration = None
    if rank == 0 and world == 1 and not args.no_iteration_leg:
        iteration = iteration_leg(ep, scene, sd, integ, wl, sensor, spp_job, g_dev, buf)

    # ---- hybrid phase 2 (it >= thres, optim.py:113-119): prb_reparam on sensor 0 with an L2 image loss
    phase2 = None
    if rank == 0 and world == 1 and args.phase2_leg:
        phase2 = phase2_leg(ep, scene, sd, wl, spp_job)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_port_throughput(sd, wl, synthetic_grad_in(res))
        except Exception as e:                                # the baseline leg must not take the bench line down
            cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        out = {"metric": "gradient path samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
               "roofline": roof, "cpu_baseline": cpu, "full_replay": full_leg, "weak": weak_leg, "iteration": iteration, "phase2_iteration": phase2,
               "scene": {"triangles": int(stats["n_triangles"]), "vertices": int(stats["n_vertices"]), "bvh_nodes": int(stats["n_nodes"]),
                         "bvh_build_ms": stats["build_ms"]},
               "wall_s_timed_region": t_wall, "ms_steps": [round(float(x), 3) for x in ms_steps]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
