#!/usr/bin/env python
"""Reproducible CPU timing of the SoftRas oracle (TEST / BASELINE INFRASTRUCTURE, never product code).

bench.py runs this file as a SUBPROCESS for its `cpu_baseline` leg and for `--impl reference`
(the reference has no CPU implementation -- SURVEY.md F1 -- so its CUDA kernels restated in C,
oracle/softras_oracle.c, stand in: kind "port").  A separate process because the OpenMP runtime
reads its binding environment once, at load time:

  * one hardware thread per PHYSICAL core of the affinity mask (SMT siblings dropped; read from
    /sys/devices/system/cpu/cpu*/topology), the process pinned to exactly those CPUs,
  * OMP_NUM_THREADS = that count, OMP_PROC_BIND=close, OMP_PLACES=cores, OMP_DYNAMIC=false,
    so threads never migrate and never share a core -- round 1 ran 128 unbound threads on a
    2-socket box and the same binary measured 0.29 and 1.48 frames/s on two identical boxes.

A "step" times fwd+bwd of ONE image of the workload: the whole image when that fits the
per-step budget, else a bounded sample of evenly spaced rows scaled by the whole / sample ratio
calibrated once on a whole image.  Prints one JSON object.
"""
import argparse
import importlib.util
import json
import math
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CHILD_FLAG = "B200R_CPU_BENCH_CHILD"


def physical_cpus():
    """One logical CPU per (package, core) among the CPUs this process may run on."""
    allowed = sorted(os.sched_getaffinity(0))
    seen, pick = set(), []
    for c in allowed:
        base = "/sys/devices/system/cpu/cpu%d/topology/" % c
        try:
            key = (open(base + "physical_package_id").read().strip(), open(base + "core_id").read().strip())
        except OSError:
            key = ("?", str(c))
        if key not in seen:
            seen.add(key)
            pick.append(c)
    return pick, len(allowed)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def reexec_pinned(threads):
    cpus, logical = physical_cpus()
    if threads > 0:
        cpus = cpus[:threads]
    os.sched_setaffinity(0, cpus)
    env = dict(os.environ)
    env.update({CHILD_FLAG: "1", "OMP_NUM_THREADS": str(len(cpus)), "OMP_PROC_BIND": "close", "OMP_PLACES": "cores",
                "OMP_DYNAMIC": "false", "OMP_WAIT_POLICY": "active", "B200R_LOGICAL_CPUS": str(logical)})
    os.execve(sys.executable, [sys.executable] + sys.argv, env)


def load_workloads():
    """jrender_b200/workloads.py is numpy-only; load it by path so this process never imports torch."""
    spec = importlib.util.spec_from_file_location("b200r_workloads", os.path.join(ROOT, "jrender_b200", "workloads.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--faces", type=int, default=39200)
    ap.add_argument("--image-size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--step-seconds", type=float, default=15.0, help="per-step budget; a whole image is timed when it fits")
    ap.add_argument("--threads", type=int, default=0, help="0 = every physical core of the affinity mask")
    ap.add_argument("--one-thread-seconds", type=float, default=4.0, help="budget of the extra 1-thread measurement (0 = skip)")
    args = ap.parse_args()
    if os.environ.get(CHILD_FLAG) != "1":
        reexec_pinned(args.threads)

    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import softras as osr
    wl = load_workloads()
    cores = len(os.sched_getaffinity(0))
    nf, H = args.faces, args.image_size
    fv, tex = wl.make_scene(nf, batch=1)
    P = osr.Params(image_size=H)
    g = np.random.default_rng(2).uniform(-1.0, 1.0, (1, 4, H, H)).astype(np.float32)

    def run(stride, nthreads):
        t0 = time.perf_counter()
        out = osr.forward(fv, tex, P, row_stride=stride, nthreads=nthreads)
        t1 = time.perf_counter()
        osr.backward(fv, tex, out, g, P, row_stride=stride, nthreads=nthreads)
        return t1 - t0, time.perf_counter() - t1

    run(max(1, H // 4), cores)             # page faults, OpenMP pool
    f0, b0 = run(1, cores)                 # one WHOLE image
    whole = f0 + b0 <= args.step_seconds
    stride, kf, kb = 1, 1.0, 1.0
    if not whole:
        stride = int(min(H // 8, max(2, math.ceil((f0 + b0) / max(args.step_seconds, 1e-3)))))
        fs, bs = run(stride, cores)
        kf, kb = f0 / max(fs, 1e-9), b0 / max(bs, 1e-9)   # whole / sample, calibrated once
    steps = []
    for _ in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        fa, ba = run(stride, cores)
        steps.append({"wall_s": time.perf_counter() - t0, "image_s": fa * kf + ba * kb, "fwd_s": fa * kf, "bwd_s": ba * kb})
    timed = steps[args.warmup:]
    image_s = float(np.mean([s["image_s"] for s in timed]))
    rows = len(range(0, H, stride))
    out = {
        "frames_per_s": 1.0 / image_s, "image_s": image_s, "fwd_s": float(np.mean([s["fwd_s"] for s in timed])),
        "bwd_s": float(np.mean([s["bwd_s"] for s in timed])), "step_wall_s": [s["wall_s"] for s in timed],
        "threads": cores, "logical_cpus_in_mask": int(os.environ.get("B200R_LOGICAL_CPUS", cores)), "cpu_model": cpu_model(),
        "binding": "one thread per physical core, OMP_PROC_BIND=close OMP_PLACES=cores, process pinned",
        "whole_image_first_s": f0 + b0, "rows_per_step": rows, "image_size": H, "faces": nf,
        "sample": ("every step timed the WHOLE %dx%d image (%d faces) fwd+bwd, no extrapolation" % (H, H, nf)) if whole else
                  ("every step timed %d evenly spaced rows of one %dx%d image (%d faces) fwd+bwd, scaled by the whole-image / "
                   "sample ratio calibrated once on a whole image" % (rows, H, H, nf)),
    }
    if args.one_thread_seconds > 0:
        # single-thread figure on a row sample sized from the all-core time (assumes at most linear speed-up)
        est = (f0 + b0) * cores
        s1 = int(min(H // 4, max(1, math.ceil(est / args.one_thread_seconds))))
        fa, ba = run(s1, 1)
        n1 = len(range(0, H, s1))
        out["one_thread"] = {"frames_per_s": 1.0 / ((fa + ba) * H / n1), "rows_timed": n1,
                             "note": "1 thread, %d evenly spaced rows x H/rows (per-image fixed costs counted H/rows times: lower bound)" % n1}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
