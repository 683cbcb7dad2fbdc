#!/usr/bin/env python3
"""bench.py — LDE+Merkle commit throughput of the B200 backend (BASELINE.json configs[2]).

One "step" = one `PolynomialBatch::from_values` of a synthetic 2^16-row x 135-column Goldilocks
trace (rate_bits 3, Poseidon Merkle cap_height 4): IFFT -> coset LDE -> Poseidon leaf hashing ->
Merkle levels -> cap. `value` times it with the trace already resident in HBM; `e2e` times the same
call through the host-facing C ABI entry point (pinned host trace in, cap out, copies inside the
timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); every rank commits its own trace (independent
proofs shard one per GPU, no data-path collective => "weak" scaling).

`--impl reference` times the CPU implementation on the host cores. The reference's own prover is
Rust over un-vendored crates and cannot be built in this image, so this arm runs the
oracle port (oracle/), multi-threaded, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "qp-zk-circuits-rm_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

DEGREE_BITS, NCOLS, RATE_BITS, CAP_HEIGHT = 16, 135, 3, 4
METRIC = "polynomial-batch commits/s (LDE+Merkle commit, 2^16 rows x 135 cols, rate_bits=3, cap_height=4)"
UNIT = "commits/s"
WORKLOAD = "PolynomialBatch::from_values 2^16 x 135 Goldilocks, rate_bits=3, Poseidon Merkle cap_height=4"
NROT = 4  # distinct input traces rotated between steps (4 x 70.8 MB > L2)


def algorithmic_counts(k=DEGREE_BITS, c=NCOLS, s=0, r=RATE_BITS, h=CAP_HEIGHT):
    """SURVEY.md §8(d) / BASELINE.md §4 per-commit work."""
    n, N = 1 << k, 1 << (k + r)
    ntt_bytes = 8 * n * c + 8 * N * (c + s)
    digest_bytes = 64 * (N - (1 << h))
    perms = N * ((c + s + 7) // 8) + (N - (1 << h))
    return dict(ntt_bytes=ntt_bytes, digest_bytes=digest_bytes, perms=perms, mults=perms * 6612)


def splitmix_trace(seed, ncols, n):
    """Seeded uniform field elements (vectorised SplitMix64, rejection of values >= p)."""
    P = 0xFFFFFFFF00000001
    total = ncols * n
    out = np.empty(0, np.uint64)
    base = np.uint64(seed)
    ctr = 0
    with np.errstate(over="ignore"):
        while out.size < total:
            m = total - out.size + 1024
            idx = np.arange(ctr + 1, ctr + 1 + m, dtype=np.uint64)
            ctr += m
            z = base + idx * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out = np.concatenate([out, z[z < np.uint64(P)]])
    return out[:total].reshape(ncols, n)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_commit_sample(threads, budget_s):
    """Time the oracle port on a bounded sample: pick the largest degree whose commit fits the budget."""
    from oracle import oracle as orc

    # calibrate Poseidon speed on this host
    small = splitmix_trace(1, NCOLS, 1 << 8)
    t0 = time.perf_counter()
    orc.batch_commit(small, RATE_BITS, CAP_HEIGHT, threads=threads, want_leaves=False, want_digests=False)
    dt = time.perf_counter() - t0
    per_row = dt / (1 << 8)
    k = DEGREE_BITS
    while k > 8 and per_row * (1 << k) > budget_s:
        k -= 1
    trace = splitmix_trace(0x5EED0001, NCOLS, 1 << k)
    return k, trace


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle as orc

    threads = os.cpu_count() or 1
    steps, warmup = args.steps, args.warmup
    budget = max(2.0, 150.0 / max(1, steps + warmup))
    k, trace = cpu_commit_sample(threads, budget)
    for _ in range(warmup):
        orc.batch_commit(trace, RATE_BITS, CAP_HEIGHT, threads=threads, want_leaves=False, want_digests=False)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.batch_commit(trace, RATE_BITS, CAP_HEIGHT, threads=threads, want_leaves=False, want_digests=False)
    dt = (time.perf_counter() - t0) / steps
    frac = float(1 << k) / float(1 << DEGREE_BITS)
    value = frac / dt  # equivalent full-size commits per second
    sample = ("one commit of 2^%d rows x %d cols per step (%.4f of the 2^%d-row workload; value scaled to "
              "full-size commits/s)" % (k, NCOLS, frac, DEGREE_BITS))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3 / frac, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Goldilocks)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU oracle port of the qp-plonky2 algorithms (the Rust "
                   "reference cannot be built here: no cargo, crates un-vendored)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_gpu(args, rank, local_rank, world):
    import torch
    import qpzk

    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctx = qpzk.Context(local_rank)
    n = 1 << DEGREE_BITS
    alg = algorithmic_counts()

    # synthetic traces: NROT distinct host traces (pinned) and their device-resident copies
    pinned = [qpzk.PinnedBuffer((NCOLS, n)) for _ in range(NROT)]
    dev = []
    for i, pb in enumerate(pinned):
        pb.array[...] = splitmix_trace(0x5EED0001 + 977 * rank + i, NCOLS, n)
        d = ctx.dev_alloc(pb.array.nbytes)
        ctx.h2d(d, pb.array)
        dev.append(d)
    cap_host = np.zeros((1 << CAP_HEIGHT, 4), np.uint64)

    ext = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        b = qpzk.PolynomialBatch.from_values_dev(ctx, dev[i % NROT], NCOLS, n, RATE_BITS, CAP_HEIGHT)
        st = ctx.stage_ms()
        b.free()
        return st

    def step_e2e(i):
        b = qpzk.PolynomialBatch.from_values(ctx, pinned[i % NROT].array, RATE_BITS, CAP_HEIGHT)
        cap = b.cap  # device -> host read of the step's result
        b.free()
        return cap

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
        acc = []
        for i in range(steps):
            acc.append(fn(i))
        with torch.cuda.stream(ext):
            e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, acc

    for i in range(args.warmup):
        step_resident(i)
    for i in range(max(1, args.warmup // 2)):
        step_e2e(i)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ms_res, stages = timed(step_resident, args.steps)
    launches = ctx.launch_count() - l0
    ms_e2e, caps = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # IMAD peak for the Poseidon roofline (dependency-free mad.wide.u32 / mad.lo.u32 loops)
    imad_wide = ctx.measure_imad_peak(1)
    imad_lo = ctx.measure_imad_peak(0)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        threads = os.cpu_count() or 1
        k, trace = cpu_commit_sample(threads, 20.0)
        t0 = time.perf_counter()
        want = orc.batch_commit(trace, RATE_BITS, CAP_HEIGHT, threads=threads, want_leaves=False,
                                want_digests=False)
        dt = time.perf_counter() - t0
        frac = float(1 << k) / float(n)
        cpu = {"value": frac / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "one oracle commit of 2^%d rows x %d cols, %d threads, %.2f s (%.4f of the workload; "
                         "scaled to full-size commits/s)" % (k, NCOLS, threads, dt, frac)}
        # the timed GPU path must agree with the oracle on that very sample
        b = qpzk.PolynomialBatch.from_values(ctx, trace, RATE_BITS, CAP_HEIGHT)
        if not np.array_equal(b.cap, want["cap"]):
            raise SystemExit("bench: GPU cap != oracle cap on the CPU-baseline sample")
        b.free()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        avg = {k2: float(np.mean([s[k2] for s in stages])) for k2 in stages[0]}
        ntt_ms = avg["ifft"] + avg["lde"]
        hash_ms = avg["leaf_hash"] + avg["merkle_levels"]
        achieved = alg["ntt_bytes"] / (ntt_ms * 1e-3) / 1e9
        peak = float(peaks["hbm_gbs"])
        total_commits = args.steps * world
        line = {
            "metric": METRIC, "value": total_commits / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Goldilocks)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "degree_bits": DEGREE_BITS, "ncols": NCOLS, "rate_bits": RATE_BITS,
                       "cap_height": CAP_HEIGHT, "blinding": False,
                       "l2": "%d distinct %.1f MB traces rotated between steps; each step streams a 566 MB LDE "
                             "(> 126 MB L2)" % (NROT, 8 * NCOLS * n / 1e6),
                       "parallelism": "independent commits, one per GPU" if world > 1 else "1 GPU"},
            "e2e": {"value": total_commits / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": 8 * NCOLS * n, "d2h_bytes_per_step": int(cap_host.nbytes)},
            "gpu_launches": int(launches),
            "stage_ms": avg,
            "roofline": {"bound": "hbm", "kernel": "IFFT + coset-LDE NTT passes (k_ntt_pass_a / k_ntt_pass_b_*)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes": alg["ntt_bytes"], "stage_ms": ntt_ms},
            "roofline_int": {"bound": "int32-multiply", "kernel": "k_leaf_hash + k_merkle_level (Poseidon)",
                             "achieved": alg["mults"] / (hash_ms * 1e-3) / 1e12, "peak": imad_wide / 1e12,
                             "unit": "T mul32/s", "frac": alg["mults"] / (hash_ms * 1e-3) / imad_wide,
                             "peak_source": "measured here: dependency-free mad.wide.u32 loop",
                             "peak_mad_lo_u32": imad_lo / 1e12, "permutations": alg["perms"],
                             "perms_per_s": alg["perms"] / (hash_ms * 1e-3), "stage_ms": hash_ms},
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)

    for pb, d in zip(pinned, dev):
        ctx.dev_free(d)
        pb.free()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # not launched by torchrun: re-launch ourselves with one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
