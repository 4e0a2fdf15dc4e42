#!/usr/bin/env python
"""bench.py -- acquisition steps/sec of the CODA hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 50 --warmup 5                 # our arm, one JSON line
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W  # N-axis sharded over N GPUs
    python bench.py --impl reference --steps 2 --warmup 1          # reference algorithm on the host cores

One step = get_next_item_to_label() -> oracle(idx) -> add_label() -> get_best_model_prediction()
(reference main.py:91-94).  Workload: synthetic M=256, N=1e6, C=100 (BASELINE.json configs[2]),
strong scaling: the N axis is split over the ranks.

  value  steps/s of the host-free device loop (labels resident in HBM; pick = arg-max, first index on equal values),
         CUDA-event timed, max over ranks;
  e2e    steps/s through the public ``coda_b200.CODA`` API with a HOST oracle: per step a pinned
         H2D copy of {idx, class} and a D2H read of the selection report and the best-model index;
  roofline  the dominant kernel of the timed region, CUDA events on the launching stream;
  cpu_baseline  the oracle (CPU restatement of coda/coda.py) on a bounded sample, extrapolated.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg3": dict(H=256, N=1_000_000, C=100),     # BASELINE.json configs[2] / configs[3]
    "cfg2": dict(H=64, N=50_000, C=10),          # BASELINE.json configs[1] (parity config)
    "mini": dict(H=32, N=20_000, C=10),          # smoke-sized
}
METRIC = "acquisition steps/sec (M=256,N=1e6,C=100)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="incremental", choices=["incremental", "recompute", "recompute_all"])
    ap.add_argument("--extra-modes", default="recompute", help="comma list of other modes to time briefly ('' = none)")
    ap.add_argument("--extra-steps", type=int, default=5)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--dense", action="store_true", help="worst-case synthetic slab (wrong class uniform)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def count(self):
        try:
            with open(self.f.name) as f:
                return sum(1 for _ in f)
        except Exception:
            return 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------
_CPU_SEL = {}


def cpu_baseline(wl, seconds, seed, threads=None):
    """Reference algorithm (oracle port of coda/coda.py) on the host cores, bounded sample, extrapolated.
    A full CPU step at cfg3 is ~days (6.55e12 quadrature cells), so: time whole 100-item chunks of the EIG
    loop (coda.py:262-279) on a 4096-item sub-slab for ~`seconds`, plus one update_pi_hat and one
    _prefilter on the sub-slab, and scale linearly to N items (chunks are independent and equal-cost)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import coda_oracle
    from coda_b200.synth import synth
    # the reference's ATen CPU kernels use every host core they are given; torchrun pins OMP_NUM_THREADS=1, undo that
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    torch.set_num_threads(threads or avail)
    H, N, C = wl["H"], wl["N"], wl["C"]
    n_sub = min(N, 4096)
    key = (H, N, C, seed)
    if key not in _CPU_SEL:      # the sub-slab and the oracle state are set-up, not part of any timed sample
        preds, labels = synth(H, N, C, seed, n_lo=0, n_hi=n_sub)
        _CPU_SEL[key] = coda_oracle.OracleSelector(preds)
    sel = _CPU_SEL[key]
    t_chunks, n_items = 0.0, 0
    t0 = time.perf_counter()
    cand = sel.candidates()
    t_pref = time.perf_counter() - t0
    # the reference's loop body handles 100 items per chunk (coda.py:235); when the time budget of one sample is
    # short the sample is a smaller batch of the same loop body (cost is linear in the items of a batch)
    bs = coda_oracle.CHUNK if seconds >= 15 else max(4, min(coda_oracle.CHUNK, int(4 * seconds)))
    k = 0
    while (t_chunks < seconds or k == 0) and (k + 1) * bs <= len(cand):
        ids = cand[k * bs:(k + 1) * bs]
        t0 = time.perf_counter()
        sel.eig_scores(ids, chunk=bs)
        t_chunks += time.perf_counter() - t0
        n_items += len(ids)
        k += 1
    t0 = time.perf_counter()
    coda_oracle.consensus_marginals(sel.dirichlets, sel.preds)
    t_pi = time.perf_counter() - t0
    t0 = time.perf_counter()
    sel.get_pbest()
    t_pb = time.perf_counter() - t0
    frac_cand = len(cand) / n_sub
    step_s = (t_chunks / n_items) * (N * frac_cand) + (t_pi + t_pref) * (N / n_sub) + 2 * t_pb
    cells = n_items * C * H * coda_oracle.QUAD_NODES
    return dict(value=1.0 / step_s, unit="steps/s", cores=torch.get_num_threads(), kind="port",
                sample=(f"extrapolated: {k} chunks x {bs} items of the EIG loop ({t_chunks:.1f}s, {cells / t_chunks:.3g} cells/s) "
                        f"+ update_pi_hat + prefilter on a {n_sub}-item sub-slab, scaled to N={N}"),
                step_seconds=step_s)


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python reference cannot travel to
    the GPU box) on this box's host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    per = max(1.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))   # whole run: a few minutes
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(wl, per, args.seed)
        if i >= args.warmup:
            vals.append(r)
    step_s = statistics.mean(v["step_seconds"] for v in vals)
    v = 1.0 / step_s
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic M={wl['H']} N={wl['N']} C={wl['C']} ({args.workload})", "mode": "reference-cpu"},
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": vals[-1]["cores"], "kind": "port", "sample": vals[-1]["sample"]},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    # stdout carries exactly one JSON line: park the real stdout and point fd 1 at stderr while libraries
    # (NCCL's version banner, torch warnings) may write
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from coda_b200 import CODA, SyntheticDataset
    from coda_b200.dist import LocalComm, TorchComm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
        comm = TorchComm()
    else:
        comm = LocalComm()
    wl = WORKLOADS[args.workload]
    H, N, C = wl["H"], wl["N"], wl["C"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")

    t0 = time.time()
    ds = SyntheticDataset(H, N, C, seed=args.seed, device=dev, dense=args.dense, rank=rank, world=world)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    labels_dev = ds.labels.to(dev)
    labels_host = ds.labels_host.numpy()

    def make(mode):
        random.seed(0)
        t = time.time()
        s = CODA(ds, mode=mode, comm=comm)
        torch.cuda.synchronize()
        return s, time.time() - t

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def device_loop(sel, warm, steps, profile_only=None):
        eng = sel.engine
        hist_idx = torch.zeros(warm + steps, dtype=torch.int64, device=dev)
        hist_q = torch.zeros(warm + steps, dtype=torch.float32, device=dev)
        for k in range(warm):
            eng.device_step(labels_dev, k, hist_idx, hist_q)
        barrier()
        launches0 = eng.counters["launches"]
        if profile_only is not None:
            eng.start_profile(profile_only or None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(warm, warm + steps):
            eng.device_step(labels_dev, k, hist_idx, hist_q)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        prof = eng.stop_profile() if profile_only is not None else {}
        eng.check_flags(sync=True)
        return ms, eng.counters["launches"] - launches0, prof, hist_idx.cpu().tolist()

    def api_loop(sel, warm, steps):
        """main.py:91-94 with a host oracle; every step copies {idx, class} H2D from pinned memory and reads
        the selection report + best model back."""
        best_host = torch.zeros(1, dtype=torch.int64).pin_memory()
        picks = []

        def one():
            idx, q = sel.get_next_item_to_label()
            t = int(labels_host[idx])                          # oracle(idx), host-resident labels
            sel.add_label(idx, t, q)
            b = sel.get_best_model_prediction()
            best_host.copy_(b.view(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            picks.append((idx, int(best_host[0])))
        for _ in range(warm):
            one()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = max_over_ranks(max(e0.elapsed_time(e1), wall * 1e3))
        return ms, picks

    # ---- our arm -------------------------------------------------------------------------------------
    sel, t_init = make(args.mode)
    eng = sel.engine
    sampler = ClockSampler(local_rank) if rank == 0 else None
    hot = ["coda_b200_pi_rank1", "coda_b200_pair_gain", "coda_b200_pair_rows", "coda_b200_pair_rows_tc",
           "coda_b200_eig_points", "coda_b200_pi_full"]
    ms, launches, prof, picks_dev = device_loop(sel, args.warmup, args.steps, profile_only=hot)
    value = args.steps / (ms / 1e3)

    # per-kernel shares over a few fully instrumented steps (not part of `value`)
    _, _, prof_all, _ = device_loop(sel, 0, min(20, args.steps), profile_only=[])

    e2e_steps = args.e2e_steps or min(args.steps, 200)
    ms_e2e, picks_api = api_loop(sel, max(1, min(args.warmup, 3)), e2e_steps)
    # the clock sampler has been running since before the warm-up; a very short run may end before nvidia-smi has
    # produced samples, so keep the same load on (untimed) until a few exist
    t_wait = time.time()
    while True:
        more = torch.tensor([1 if (sampler is not None and sampler.count() < 5 and time.time() - t_wait < 3.0) else 0],
                            device=dev)
        if world > 1:
            dist.broadcast(more, src=0)        # every rank runs the same number of (collective) extra steps
        if not int(more.item()):
            break
        device_loop(sel, 0, 20)
    clocks = sampler.stop() if sampler else {}
    e2e = e2e_steps / (ms_e2e / 1e3)
    h2d = eng.sel_host.numel() * 8
    d2h = eng.rep_host.numel() * 8 * (world if world > 1 else 1) + 8

    extra = {}
    for m in [x for x in args.extra_modes.split(",") if x and x != args.mode]:
        del sel, eng
        torch.cuda.empty_cache()
        sel, t_i = make(m)
        eng = sel.engine
        ms_m, _, prof_m, _ = device_loop(sel, 2, args.extra_steps, profile_only=hot)
        extra[m] = {"value": args.extra_steps / (ms_m / 1e3), "unit": "steps/s", "ms_per_step": ms_m / args.extra_steps,
                    "init_s": t_i, "kernel_ms": {k.replace("coda_b200_", ""): v[1] / max(1, v[0]) for k, v in prof_m.items()}}

    # ---- roofline of the dominant kernel of the timed region -----------------------------------------
    n_loc = eng.N
    npairs, Hp = eng.npairs, eng.Hp
    alg_bytes = {   # algorithmic bytes per launch, per rank
        "coda_b200_pi_rank1": 4 * H * n_loc + 4 * n_loc * C + 4 * n_loc,
        "coda_b200_pair_gain": 4 * npairs * Hp + 4 * npairs,
        "coda_b200_eig_points": 4 * n_loc * C + 4 * n_loc + 6 * eng.n_entries,
        "coda_b200_pi_full": 4 * H * n_loc * C + 4 * n_loc * C,
    }
    roof = None
    if prof:
        dom = max(prof, key=lambda k: prof[k][1])
        cnt, tot = prof[dom]
        avg_ms = tot / max(1, cnt)
        if dom in alg_bytes:
            ach = alg_bytes[dom] / (avg_ms * 1e-3) / 1e9
            roof = {"kernel": dom.replace("coda_b200_", "k_"), "bound": "hbm", "achieved": ach, "peak": hbm_peak,
                    "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_src,
                    "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": alg_bytes[dom],
                    "share_of_step": tot / ms}
        elif dom == "coda_b200_pair_rows_tc":
            # 9 bf16 MMAs of 128 x 256 x Hp (3 dL limbs) / 128 x Hp x 256 (2 tables x 3 cross terms) per 128-pair tile
            tiles_per_launch = eng.ntiles if args.mode != "incremental" else max(1, eng.ntiles // C)
            flops = 9 * 2 * 128 * 256 * eng.Hp * tiles_per_launch
            tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
            ach = flops / (avg_ms * 1e-3) / 1e12
            roof = {"kernel": "k_pair_rows_tc", "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": ach / tf_peak, "traffic": None, "peak_source": peak_src, "avg_launch_ms": avg_ms,
                    "algorithmic_flops_per_launch": flops, "share_of_step": tot / ms}
        else:
            roof = {"kernel": dom.replace("coda_b200_", "k_"), "bound": "hbm", "avg_launch_ms": avg_ms,
                    "share_of_step": tot / ms, "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None,
                    "traffic": None}

    if roof is not None:
        try:    # measured DRAM traffic of the same kernel/config from the committed ncu capture (not measurable live)
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")))
            roof["traffic"] = tr.get(f"{args.workload}/{world}/{args.mode}", {}).get(roof["kernel"])
            roof["traffic_source"] = "profiles/r1_step_kernels_ncu.txt" if roof["traffic"] else None
        except Exception:
            pass
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # rank 0, N=1 only (the reference arm covers N>1)
        cpu = cpu_baseline(wl, args.cpu_seconds, args.seed)
        cpu.pop("step_seconds", None)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"synthetic M={H} N={N} C={C} ({args.workload}{', dense' if args.dense else ''}), N-axis sharded over {world} GPU(s)",
                "mode": args.mode, "l2": "per-step working set (slab gather + row cache + U) >> 126 MB L2; no flush needed",
                "tie_rule_value": "arg-max, first index (device loop)", "tie_rule_e2e": "random.choice (coda.py:308)",
                "pairs": npairs, "heavy_pairs": eng.n_heavy, "tensor_core_rows": bool(eng.use_tc), "entries_per_item": eng.n_entries / max(1, n_loc),
                "gen_s": t_gen, "init_s": t_init, "shadow_models": eng.n_shadow,
            },
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "steps/s", "ms_per_step": ms_e2e / e2e_steps, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": launches,
            "roofline": roof,
            "kernel_ms": {k.replace("coda_b200_", ""): v[1] / max(1, v[0]) for k, v in prof_all.items()},
            "modes": extra,
            "cpu_baseline": cpu,
            "first_picks": {"device_loop": picks_dev[:8], "api": [p[0] for p in picks_api[:8]]},
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
