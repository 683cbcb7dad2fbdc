#!/usr/bin/env python3
"""bench.py — wormhole proofs/s and LDE+Merkle commit ms of the B200 backend.

Headline workload (BASELINE.json `metric`: "wormhole proofs/s; LDE+Merkle commit ms @2^k rows"):
a "step" is one wormhole proof of the bench-data shape (the ZK circuit of
wormhole/bench-data — 2^14 rows x 135 wires, 80 routed, 6 gates, rate_bits 3, cap_height 4, FRI
arities [4,4,4], 28 queries, 16 PoW bits, 4 salt columns per blinded oracle) on a synthetic
SATISFYING witness: commit wires -> Z/partial products -> commit -> quotient -> commit -> openings
-> FRI (combine, fold+commit, PoW, queries) -> ProofWithPublicInputs bytes. Proofs are independent,
so they shard one per GPU stream (BASELINE configs[3]); `--streams` proofs are in flight per GPU.
  value : witness matrices already resident in HBM when the clock starts
  e2e   : the public entry point with HOST buffers (pinned witness in, proof bytes out)
The same run also times the PolynomialBatch commit microbench (configs[2]: 2^16 x 135) and reports
its stage times with the two rooflines SURVEY.md §8(d) asks for: NTT stages against measured HBM
bandwidth (`roofline`) and Poseidon hashing against the measured integer multiply-add peak
(`roofline_int`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--impl reference]

N > 1 is launched by torchrun, one rank per GPU, every rank proving its own witnesses (weak scaling,
no data-path collective). `--impl reference` times the CPU path on the host cores: the reference's
Rust prover cannot be built in this image (no cargo, crates un-vendored - DESIGN.md), so that arm
runs the oracle port (oracle/prover.hpp) with all host threads.
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

PROOF_K, PROOF_ZK = 14, True
SALT_SEED = np.array([0x5EED0003, 0x0123456789ABCDEF, 0xFEDCBA9876543210, 0x0F1E2D3C4B5A6978], np.uint64)
DEGREE_BITS, NCOLS, RATE_BITS, CAP_HEIGHT = 16, 135, 3, 4
METRIC = "wormhole proofs/s (bench-data shape: 2^14 rows x 135 wires, ZK, rate_bits=3, cap_height=4, 28 queries)"
UNIT = "proofs/s"
WORKLOAD = ("wormhole single-proof generation, bench-data shape (2^14 x 135 wires, 80 routed, gates "
            "Noop/Constant/PublicInput/BaseSum63/Arithmetic20/Poseidon, FRI [4,4,4], salted), synthetic satisfying witness")


# dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture per kernel
# IFFT (two passes): profiles/r1_ntt_v2_ncu_full.txt k_ntt_pass_a<0>, k_ntt_pass_b_transpose;
# LDE (one pass, cluster kernel): profiles/r2_ntt_cluster_2p16_ncu_full.txt
NTT_DRAM_BYTES_PER_COMMIT = int((71.430912 + 28.294144 + 70.869504 + 25.209344 + 75.771648 + 512.583936) * 1e6)
LEAF_HASH_DRAM_BYTES_PER_COMMIT = int((566.815232 + 17.787392) * 1e6)   # profiles/r1_leaf_hash_v5_twoplane_ncu_full.txt


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
    """nvidia-smi clocks / throttle reasons. The sampler runs from before the warm-up (nvidia-smi needs a
    few hundred ms to print its first row); only rows stamped inside [mark_begin, mark_end] - the timed
    region - are summarised."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def wait_first(self, timeout=5.0):
        t = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t < timeout:
            time.sleep(0.02)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if self.t0 is not None and not (self.t0 <= ts <= (self.t1 or ts) + 0.06):
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "samples_total": len(self.rows),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
class OracleProvider:
    """Hashes for the synthetic-witness generator when no GPU is involved (--impl reference)."""

    def __init__(self):
        from oracle import oracle as orc
        self.poseidon_tables = orc.poseidon_tables
        self.hash_no_pad = orc.hash_no_pad


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle as orc
    from qpzk import synth

    threads = os.cpu_count() or 1
    prov = OracleProvider()

    def make(k):
        circ = synth.build(k, zk=PROOF_ZK, seed=1, provider=prov)
        return circ, orc.Circuit(circ["common"], circ["digest"], circ["constants_sigmas"], threads=threads)

    # bounded sample: the largest degree whose (steps + warmup) proofs fit ~150 s on this host
    circ, oc = make(10)
    t0 = time.perf_counter()
    oc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    t10 = time.perf_counter() - t0
    k = PROOF_K
    while k > 10 and t10 * (1 << (k - 10)) * (args.steps + args.warmup) > 150.0:
        k -= 1
    if k != 10:
        circ, oc = make(k)
    for _ in range(args.warmup):
        oc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        proof = oc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    rc, _ = orc.verify(circ["common"], oc.verifier_only_bytes(), proof)
    if rc != 0:
        raise SystemExit("reference arm: oracle proof not accepted by the oracle verifier")
    # the port's hashing rate, for scale: MerkleTree::new over 2^15 rows of 139 elements (19 permutations per row)
    rng = np.random.default_rng(1)
    lv = rng.integers(0, 0xFFFFFFFF00000001, size=(1 << 15, 139), dtype=np.uint64)
    t0 = time.perf_counter()
    orc.merkle_new(lv, 4, threads=threads)
    mperm = (1 << 15) * 19 / (time.perf_counter() - t0) / 1e6
    frac = float(1 << k) / float(1 << PROOF_K)
    dt = dt / frac          # scaled to the full 2^14-row proof (prover work is ~linear in rows)
    value = 1.0 / dt
    sample = ("one full proof per step" if k == PROOF_K else
              "one proof of the same circuit family at 2^%d rows per step (%.4f of the 2^%d-row workload; value "
              "scaled linearly in rows)" % (k, frac, PROOF_K))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Goldilocks)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU oracle port of qp-plonky2's prove() (the Rust reference cannot be "
                   "built here: no cargo, crates un-vendored); C++, %d host threads, AVX-512 where the CPU has it (8-way "
                   "Poseidon for Merkle hashing and proof of work, vectorised NTT butterflies), scalar otherwise" % threads},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "avx512": bool(orc.have_avx512()), "poseidon_Mperm_per_s": mperm},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def commit_microbench(ctx, qpzk, steps, warmup, rank, blinding=False):
    """BASELINE configs[2]: per-stage device times of PolynomialBatch::from_values, 2^16 x 135; with
    blinding = True, 4 salt columns from SplitMix64 seed 0x5eed0002 join every leaf."""
    n = 1 << DEGREE_BITS
    nrot = 4  # 4 x 70.8 MB distinct traces > 126 MB L2
    dev = []
    for i in range(nrot):
        tr = splitmix_trace(0x5EED0001 + 977 * rank + i, NCOLS, n)
        d = ctx.dev_alloc(tr.nbytes)
        ctx.h2d(d, tr)
        dev.append(d)
    salts = None
    if blinding:
        sa = splitmix_trace(0x5EED0002, 4, n << RATE_BITS)
        sp = ctx.dev_alloc(sa.nbytes)
        ctx.h2d(sp, sa)
        salts = (sp, 4)
    stages = []
    for i in range(warmup + steps):
        b = qpzk.PolynomialBatch.from_values_dev(ctx, dev[i % nrot], NCOLS, n, RATE_BITS, CAP_HEIGHT, salts=salts)
        st = ctx.stage_ms()
        b.free()
        if i >= warmup:
            stages.append(st)
    for d in dev:
        ctx.dev_free(d)
    if salts:
        ctx.dev_free(salts[0])
    return {k2: float(np.mean([s[k2] for s in stages])) for k2 in stages[0]}


def commit_sweep(ctx, qpzk, ks=(12, 13, 14, 15, 16, 17, 18, 19, 20)):
    """The metric's "LDE+Merkle commit ms @ 2^k rows" curve: PolynomialBatch::from_values of a
    2^k x 135 trace, rate_bits 3, cap_height 4, device-resident input, CUDA-event stage times (best of 3)."""
    out = {}
    for k in ks:
        n = 1 << k
        tr = splitmix_trace(0x5EED0001 + k, NCOLS, n)
        d = ctx.dev_alloc(tr.nbytes)
        ctx.h2d(d, tr)
        best = None
        for _ in range(4):
            b = qpzk.PolynomialBatch.from_values_dev(ctx, d, NCOLS, n, RATE_BITS, CAP_HEIGHT)
            st = ctx.stage_ms()
            b.free()
            tot = sum(st.values())
            if best is None or tot < best[0]:
                best = (tot, st)
        ctx.dev_free(d)
        alg = algorithmic_counts(k=k)
        ntt = best[1]["ifft"] + best[1]["lde"]
        hsh = best[1]["leaf_hash"] + best[1]["merkle_levels"]
        out["2^%d" % k] = {"ms": best[0], "ntt_ms": ntt, "poseidon_ms": hsh,
                           "ntt_GBps_algorithmic": alg["ntt_bytes"] / (ntt * 1e-3) / 1e9,
                           "Mperm_per_s": alg["perms"] / (hsh * 1e-3) / 1e6}
    return out


def sharded_commit_bench(ctx, qpzk, torch, dist, rank, world, steps, warmup):
    """BASELINE configs[2] at N GPUs: ONE 2^16 x 135 commit sharded over the ranks by cap subtrees (whole
    LDE cosets), every rank holding the trace; the only exchange is the NCCL all-gather of the 16 subtree
    roots. Device time, max over ranks. The gathered cap must equal the unsharded cap."""
    from qpzk import dist as qdist
    n = 1 << DEGREE_BITS
    tr = splitmix_trace(0x5EED0001, NCOLS, n)          # same trace on every rank
    d = ctx.dev_alloc(tr.nbytes)
    ctx.h2d(d, tr)
    b0, e0 = qdist.shard_subtrees(rank, world, CAP_HEIGHT, RATE_BITS)
    times, wall, cap = [], [], None
    stream = qdist._ext_stream(ctx)
    for i in range(warmup + steps):
        dist.barrier()
        torch.cuda.synchronize()
        # device time on the context's stream, from the first kernel of the commit to the end of the all-gather; the
        # commit is enqueued without waiting, so the all-gather follows it without a host round trip
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        sh = qpzk.PolynomialBatch.from_values_shard_dev(ctx, d, NCOLS, n, RATE_BITS, CAP_HEIGHT, b0, e0, enqueue_only=True)
        qdist.allgather_cap_nccl(sh, sync=False)
        ev1.record(stream)
        ctx.sync()
        w = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([ev0.elapsed_time(ev1), w], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if i >= warmup:
            times.append(float(ms[0].item()))
            wall.append(float(ms[1].item()))
        cap = sh.cap
        sh.free()
    ok = None
    if rank == 0:
        full = qpzk.PolynomialBatch.from_values_dev(ctx, d, NCOLS, n, RATE_BITS, CAP_HEIGHT)
        ok = bool(np.array_equal(full.cap, cap))
        full.free()
        if not ok:
            raise SystemExit("bench: sharded commit cap != unsharded cap")
    ctx.dev_free(d)
    return {"n_gpus": world, "ms": float(np.mean(times)), "host_wall_ms": float(np.mean(wall)), "subtrees_per_gpu": e0 - b0,
            "exchange": "NCCL all_gather_into_tensor of %d x 32 B subtree roots, in place on the device cap"
                        % (1 << CAP_HEIGHT),
            "cap_equals_unsharded": ok}


def sharded_proof_bench(ctx, qpzk, torch, dist, rank, world, k, reps=3):
    """BASELINE configs[4] on N GPUs: ONE recursion-shaped (aggregation) proof of 2^k rows, every rank working on
    its cap subtrees (qpzk_sprove_*, NCCL exchanges on the context's stream), witness resident in HBM on every
    rank. Device time from the first enqueue to the proof bytes, max over ranks. Rank 0 also proves it alone: the
    bytes must be equal."""
    from qpzk import dist as qdist
    from qpzk import synth
    bc = synth.build_recursion(k, zk=True, seed=10, provider=synth.GpuProvider(ctx))
    circ = qpzk.Circuit(ctx, bc["common"], bc["digest"], bc["constants_sigmas"])
    dw = ctx.dev_alloc(bc["wires"].nbytes)
    ctx.h2d(dw, bc["wires"])
    ds = []
    for sa in bc["salts"]:
        p = ctx.dev_alloc(sa.nbytes)
        ctx.h2d(p, sa)
        ds.append(p)
    times, proof = [], None
    for i in range(reps + 1):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        proof = qdist.prove_sharded_nccl(circ, dw, bc["public_inputs"], ds, CAP_HEIGHT, RATE_BITS, on_device=True)
        ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if i:
            times.append(float(ms.item()))
    stages = circ.stage_ms()
    out = None
    if rank == 0:
        lat = []
        for i in range(3):
            t0 = time.perf_counter()
            alone = circ.prove_dev(dw, bc["public_inputs"], ds)
            if i:
                lat.append((time.perf_counter() - t0) * 1e3)
        if alone != proof:
            raise SystemExit("bench: %d-GPU sharded proof bytes != single-GPU proof bytes" % world)
        from oracle import oracle as orc  # the checker: the restated plonky2 verifier
        rc, _ = orc.verify(bc["common"], circ.verifier_only_bytes(), proof)
        if rc != 0:
            raise SystemExit("bench: sharded proof rejected by the restated verifier (rc %d)" % rc)
        out = {"workload": "one aggregation-shaped proof: 2^%d rows x 135 wires, 14-gate recursion set, ZK, witness "
                           "resident in HBM on every rank" % k,
               "n_gpus": world, "latency_ms_median": float(np.median(times)), "latency_ms_min": float(np.min(times)),
               "single_gpu_latency_ms": float(np.median(lat)), "speedup": float(np.median(lat) / np.median(times)),
               "bytes_equal_single_gpu_proof": True, "verifier_accepts": True, "proof_bytes": len(proof),
               "stage_ms_rank0": stages,
               "exchange": "NCCL on the context's stream: 3 all-gathers of 16 x 32 B subtree roots, all-gather of the "
                           "quotient values (2 x 2^%d x 8 B), all-reduce of 28 x 3 opened rows" % (k + 3)}
    dist.barrier()
    for p in [dw] + ds:
        ctx.dev_free(p)
    circ.free()
    return out


def tree_schedule_bench(ctxs, qpzk, torch, dist, rank, world, reps=3):
    """The reference aggregator's default tree (8 leaves, branching factor 2: 4 + 2 + 1 dependent node proofs,
    /root/reference/wormhole/aggregator/src/circuits/tree.rs:17-20,55-103) dealt over the ranks: every node is a
    2^13-row recursion-shaped proof, the nodes of a level run concurrently (one per GPU, or per stream on one
    GPU), the node proofs are all-gathered after each level. Wall time from the leaf proofs to the root proof,
    max over ranks."""
    from qpzk import aggregate as agg
    from qpzk import synth
    ak = 13
    ac = synth.build_recursion(ak, zk=True, seed=9, provider=synth.GpuProvider(ctxs[0]))
    slots = min(len(ctxs), 4)
    circs = [qpzk.Circuit(c, ac["common"], ac["digest"], ac["constants_sigmas"]) for c in ctxs[:slots]]
    pw = qpzk.PinnedBuffer(ac["wires"].shape)
    pw.array[...] = ac["wires"]
    ps = [qpzk.PinnedBuffer(sa.shape) for sa in ac["salts"]]
    for a, sa in zip(ps, ac["salts"]):
        a.array[...] = sa
    node_len = len(circs[0].prove(pw.array, ac["public_inputs"], [a.array for a in ps]))
    leaves = [bytes([i]) * node_len for i in range(8)]

    def begin(level, index, children, slot):
        # building the node circuit and its witness from `children` is host work outside this backend
        circs[slot].prove_begin(pw.array, ac["public_inputs"], [a.array for a in ps])
        return slot

    gather = agg.torch_all_gather(node_len, device=torch.device("cuda", ctxs[0].device)) if world > 1 else None
    # upper levels have fewer nodes than GPUs: the idle ranks join and a node becomes ONE proof sharded over a
    # sub-group of ranks (qpzk_sprove_*; NCCL exchanges on the context's stream, same bytes on every rank)
    groups = agg.make_rank_groups(world, rank) if world > 1 else {}

    def prove_group(level, index, children, ranks):
        from qpzk import dist as qdist
        return qdist.prove_sharded_nccl(circs[0], pw.array, ac["public_inputs"], [a.array for a in ps], CAP_HEIGHT,
                                        RATE_BITS, on_device=False, group=groups[len(ranks)])

    def run(grouped):
        times, root = [], None
        for i in range(reps + 1):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            root, levels = agg.aggregate_tree(leaves, 2, begin, lambda slot: circs[slot].prove_end(), rank, world, gather,
                                              prove_group=prove_group if grouped else None)
            ms = (time.perf_counter() - t0) * 1e3
            if dist is not None:
                t = torch.tensor([ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            if i:
                times.append(ms)
        return times, root

    times, root = run(False)
    out = {"workload": "aggregation tree of 8 leaf proofs, branching factor 2: 4 + 2 + 1 dependent node proofs of 2^%d "
                       "rows (recursion gate set, ZK, host buffers), node proofs all-gathered per level" % ak,
           "n_gpus": world, "latency_ms_median": float(np.median(times)), "latency_ms_min": float(np.min(times)),
           "node_proofs": 7, "node_proof_bytes": node_len, "root_proof_bytes": len(root)}
    if world > 1:
        gtimes, groot = run(True)
        if groot != root:
            raise SystemExit("bench: the tree with group-sharded node proofs ends in a different root proof")
        out["one_gpu_per_node_latency_ms_median"] = out["latency_ms_median"]
        out["one_gpu_per_node_latency_ms_min"] = out["latency_ms_min"]
        out["latency_ms_median"], out["latency_ms_min"] = float(np.median(gtimes)), float(np.min(gtimes))
        out["ranks_per_node_by_level"] = [agg.ranks_per_node(c, world) for c in agg.tree_levels(8, 2)]
        out["schedule"] = ("levels with fewer nodes than GPUs shard every node proof over world / nodes ranks "
                           "(qpzk_sprove_*, NCCL sub-groups); same root proof bytes as one GPU per node")
    for q in circs:
        q.free()
    return out


def run_gpu(args, rank, local_rank, world):
    import torch
    import qpzk
    from qpzk import synth

    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    S = max(1, args.streams)
    HT = max(1, min(args.host_threads, S))
    # a proof is one stream-ordered enqueue (qpzk_prove_begin) and one wait (qpzk_prove_end), so HT host
    # threads keep S proofs in flight; a waiting thread spins on a core unless threads x ranks exceed the cores
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sync = args.sync
    if sync == "auto":
        sync = "spin" if HT * max(1, world) <= cores else "yield"
    ctxs = [qpzk.Context(local_rank, blocking_sync=sync == "blocking", yield_sync=sync == "yield") for _ in range(S)]
    ctx0 = ctxs[0]

    # one synthetic wormhole-shaped circuit + witness per stream
    circ = synth.build(PROOF_K, zk=PROOF_ZK, seed=1 + rank, provider=synth.GpuProvider(ctx0))
    n = 1 << PROOF_K
    circuits = [qpzk.Circuit(c, circ["common"], circ["digest"], circ["constants_sigmas"]) for c in ctxs]
    circuit_commit_ms = sum(ctx0.stage_ms().values())
    nw = circ["wires"].shape[0]
    pinned_w, pinned_s, dev_w, dev_s = [], [], [], []
    for c in ctxs:
        pw = qpzk.PinnedBuffer((nw, n))
        pw.array[...] = circ["wires"]
        pinned_w.append(pw)
        ps = [qpzk.PinnedBuffer(s.shape) for s in circ["salts"]]
        for a, s in zip(ps, circ["salts"]):
            a.array[...] = s
        pinned_s.append(ps)
        d = c.dev_alloc(pw.array.nbytes)
        c.h2d(d, pw.array)
        dev_w.append(d)
        ds = []
        for s in circ["salts"]:
            p = c.dev_alloc(s.nbytes)
            c.h2d(p, s)
            ds.append(p)
        dev_s.append(ds)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    results = [None] * HT

    def worker(t, count, resident):
        """Host thread t keeps its streams (t, t + HT, ...) full: begin on every idle stream, then end on the
        oldest proof in flight."""
        mine = list(range(t, S, HT))
        inflight, issued, done, out = [], 0, 0, None
        while done < count:
            while issued < count and len(inflight) < len(mine):
                s = mine[issued % len(mine)]
                if resident:
                    circuits[s].prove_begin_dev(dev_w[s], circ["public_inputs"], dev_s[s])
                else:   # host witness in; the salts are drawn on the device from a seed, as the reference draws its own
                    circuits[s].prove_begin(pinned_w[s].array, circ["public_inputs"], seed=SALT_SEED)
                inflight.append(s)
                issued += 1
            out = circuits[inflight.pop(0)].prove_end()
            done += 1
        results[t] = out

    def run_round(total, resident):
        # exactly `total` proofs, spread over the HT host threads (and through them over the S streams)
        th = [threading.Thread(target=worker, args=(t, total // HT + (1 if t < total % HT else 0), resident))
              for t in range(HT)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    def timed(total, resident):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # events on the default stream bracket ALL context streams: the region begins and ends with a full
        # device synchronize, and every prove call returns only after its own stream has drained
        e0.record()
        run_round(total, resident)
        torch.cuda.synchronize()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # a step = one proof on every stream of this GPU (S proofs): the timed region is steps x S proofs per GPU
    steps_total = args.steps * S
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # W warm-up steps of each kind (at least 8 of the resident kind: at 8 GPUs the first ~0.2 s after the ranks start
    # together ran 4 % below the steady state that a longer warm-up - or the e2e region that follows - sees)
    run_round(max(args.warmup, 8) * S, True)
    run_round(args.warmup * S, False)
    if rank == 0:
        sampler.wait_first()
        sampler.mark_begin()
    l0 = sum(c.launch_count() for c in ctxs)
    ms_res = timed(steps_total, True)
    launches = sum(c.launch_count() for c in ctxs) - l0
    proof = results[0]
    ms_e2e = timed(steps_total, False)
    proof_e2e = results[0]
    if rank == 0:
        sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None

    # ---- everything that needs all ranks comes first; then ranks > 0 are released (they used to spin in a
    # barrier at ~60 % "GPU busy" while rank 0 ran its single-GPU extras) ----
    sharded = sharded_proof = None
    tree = None
    if dist is not None:
        if (1 << min(CAP_HEIGHT, RATE_BITS)) % world == 0:
            sharded = sharded_commit_bench(ctx0, qpzk, torch, dist, rank, world, 5, 3)
            if not args.no_aggregator:
                sharded_proof = sharded_proof_bench(ctx0, qpzk, torch, dist, rank, world, args.shard_k)
        if not args.no_aggregator:
            tree = tree_schedule_bench(ctxs, qpzk, torch, dist, rank, world)
        dist.barrier()
        dist.destroy_process_group()
        dist = None
        if rank != 0:
            for c, d, ds in zip(ctxs, dev_w, dev_s):
                c.dev_free(d)
                for p in ds:
                    c.dev_free(p)
            for q in circuits:
                q.free()
            for c in ctxs:
                c.close()
            return
    elif not args.no_aggregator:
        tree = tree_schedule_bench(ctxs, qpzk, torch, None, 0, 1)

    # single-proof latency and stage breakdown (one stream, nothing else in flight)
    t0 = time.perf_counter()
    circuits[0].prove_dev(dev_w[0], circ["public_inputs"], dev_s[0])
    latency_ms = (time.perf_counter() - t0) * 1e3
    proof_stages = circuits[0].stage_ms()

    # BASELINE configs[1]: voting-circuit-sized single proof (non-ZK `standard_recursion_config`,
    # /root/reference/voting/src/lib.rs:348; degree ~2^9 - SURVEY 8(d)): per-proof fixed overhead / latency
    voting = None
    if rank == 0:
        vk = 9
        vc = synth.build(vk, zk=False, seed=5, provider=synth.GpuProvider(ctx0))
        vcirc = qpzk.Circuit(ctx0, vc["common"], vc["digest"], vc["constants_sigmas"])
        vd = ctx0.dev_alloc(vc["wires"].nbytes)
        ctx0.h2d(vd, vc["wires"])
        for _ in range(3):
            vproof = vcirc.prove_dev(vd, vc["public_inputs"])
        lat = []
        for _ in range(20):
            t0 = time.perf_counter()
            vcirc.prove(vc["wires"], vc["public_inputs"])
            lat.append((time.perf_counter() - t0) * 1e3)
        voting = {"workload": "voting-circuit shape: 2^%d rows x 135 wires, non-ZK, one proof at a time through "
                              "qpzk_prove with host buffers" % vk,
                  "latency_ms_median": float(np.median(lat)), "latency_ms_min": float(np.min(lat)),
                  "proof_bytes": len(vproof), "stage_ms": vcirc.stage_ms()}
        ctx0.dev_free(vd)
        vcirc.free()

    # a prover restored from files (`new_from_files`, /root/reference/wormhole/prover/src/lib.rs:105-187)
    # uploads the serialized constants_sigmas_commitment instead of recomputing it; both ways timed, same proof bytes
    restored = None
    if rank == 0:
        blob = circuits[0].commitment_bytes()
        t_new, t_restored = [], []
        for _ in range(3):
            t0 = time.perf_counter()
            qa = qpzk.Circuit(ctx0, circ["common"], circ["digest"], circ["constants_sigmas"])
            t_new.append((time.perf_counter() - t0) * 1e3)
            qa.free()
            t0 = time.perf_counter()
            qb = qpzk.Circuit(ctx0, circ["common"], circ["digest"], commitment=blob)
            t_restored.append((time.perf_counter() - t0) * 1e3)
            if _ < 2:
                qb.free()
        same = qb.prove_dev(dev_w[0], circ["public_inputs"], dev_s[0]) == circuits[0].prove_dev(dev_w[0], circ["public_inputs"], dev_s[0])
        qb.free()
        if not same:
            raise SystemExit("bench: a circuit restored from its serialized commitment proves different bytes")
        restored = {"workload": "qpzk_circuit_create (commit 84 columns x 2^%d on the device) against "
                                "qpzk_circuit_create_from_commitment (upload the serialized PolynomialBatch from pageable "
                                "host memory, nothing recomputed)" % PROOF_K,
                    "create_ms_min": float(np.min(t_new)), "from_commitment_ms_min": float(np.min(t_restored)),
                    "commitment_bytes": len(blob), "same_proof_bytes": True}
        del blob

    # BASELINE configs[4]: one aggregation-tree node - a recursion-shaped circuit (the 14-gate set an
    # in-circuit verifier instantiates, /root/reference/wormhole/aggregator/src/circuits/tree.rs:111-136) of
    # 2^13 rows, the size SURVEY 8(d) estimates for the reference's default binary tree nodes.
    aggregator = None
    if rank == 0 and not args.no_aggregator:
        ak = 13
        ac = synth.build_recursion(ak, zk=True, seed=9, provider=synth.GpuProvider(ctx0))
        acircs = [qpzk.Circuit(c, ac["common"], ac["digest"], ac["constants_sigmas"]) for c in ctxs]
        for _ in range(3):
            aproof = acircs[0].prove(ac["wires"], ac["public_inputs"], ac["salts"])
        lat = []
        for _ in range(10):
            t0 = time.perf_counter()
            acircs[0].prove(ac["wires"], ac["public_inputs"], ac["salts"])
            lat.append((time.perf_counter() - t0) * 1e3)
        astages = acircs[0].stage_ms()
        per = 8

        def aworker(s):
            for _ in range(per):
                acircs[s].prove(ac["wires"], ac["public_inputs"], ac["salts"])

        th = [threading.Thread(target=aworker, args=(s,)) for s in range(S)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        aggregator = {"workload": "aggregation-node shape: 2^%d rows x 135 wires, 14-gate recursion set, ZK, host "
                                  "buffers through qpzk_prove" % ak,
                      "latency_ms_median": float(np.median(lat)), "proofs_per_s_%d_streams" % S: S * per / dt,
                      "proof_bytes": len(aproof), "stage_ms": astages,
                      "note": "gate definitions restated from upstream plonky2, not pinned by a reference fixture"}
        for q in acircs:
            q.free()
        # a flat 16-ary aggregation node is ~2^16 rows, BASELINE configs[4] quotes ~2^17-2^18:
        # single-proof latency at 2^16, 2^17 and 2^18 rows, each proof accepted by the restated verifier
        from oracle import oracle as orc  # the checker: the restated plonky2 verifier
        for bk in ((16, 17, 18) if world == 1 else ()):   # at N > 1 the 2^18-row proof is the sharded one above
            bc = synth.build_recursion(bk, zk=True, seed=10, provider=synth.GpuProvider(ctx0))
            bcirc = qpzk.Circuit(ctx0, bc["common"], bc["digest"], bc["constants_sigmas"])
            blat, bst = [], []
            for i in range(5):     # first run warms the memory pool up to this proof size; the other four are reported
                t0 = time.perf_counter()
                bproof = bcirc.prove(bc["wires"], bc["public_inputs"], bc["salts"])
                if i:
                    blat.append((time.perf_counter() - t0) * 1e3)
                    bst.append(bcirc.stage_ms())
            rc, _ = orc.verify(bc["common"], bcirc.verifier_only_bytes(), bproof)
            if rc != 0:
                raise SystemExit("2^%d-row recursion-shaped proof rejected by the restated verifier (rc %d)" % (bk, rc))
            aggregator["flat_node_2^%d_rows" % bk] = {
                "latency_ms_median": float(np.median(blat)), "latency_ms_all": [float(x) for x in blat],
                "proof_bytes": len(bproof), "verifier_accepts": True,
                "stage_ms_median": {kk: float(np.median([st[kk] for st in bst])) for kk in bst[0]},
                "stage_ms_max": {kk: float(np.max([st[kk] for st in bst])) for kk in bst[0]}}
            bcirc.free()
            del bc

    micro = commit_microbench(ctx0, qpzk, 5, 3, rank) if rank == 0 else None
    micro_blind = commit_microbench(ctx0, qpzk, 5, 3, rank, blinding=True) if rank == 0 else None
    sweep = commit_sweep(ctx0, qpzk) if rank == 0 else None
    imad_wide = ctx0.measure_imad_peak(1)
    imad_lo = ctx0.measure_imad_peak(0)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        threads = os.cpu_count() or 1
        oc = orc.Circuit(circ["common"], circ["digest"], circ["constants_sigmas"], threads=threads)
        t0 = time.perf_counter()
        want = oc.prove(circ["wires"], circ["public_inputs"], circ["salts"])
        dt = time.perf_counter() - t0
        cpu = {"value": 1.0 / dt, "unit": UNIT, "cores": threads, "kind": "port", "avx512": bool(orc.have_avx512()),
               "sample": "one full proof of the same circuit and witness with the oracle port (C++, %d threads, AVX-512 "
                         "Poseidon / NTT where available): %.2f s" % (threads, dt)}
        # parity gate inside the bench: the timed GPU path must emit the very bytes the oracle does,
        # and the restated verifier must accept them
        if proof != want:
            raise SystemExit("bench: GPU proof bytes != oracle proof bytes")
        # the e2e arm drew its salts on the device: the oracle gets the same salts restated on the host
        want_e2e = oc.prove(circ["wires"], circ["public_inputs"], [synth.seeded_salts(SALT_SEED, o, n << RATE_BITS)
                                                                   for o in range(3)])
        if proof_e2e != want_e2e:
            raise SystemExit("bench: GPU proof with device-drawn salts != oracle proof with the restated salts")
        rc, _ = orc.verify(circ["common"], circuits[0].verifier_only_bytes(), proof)
        if rc != 0:
            raise SystemExit("bench: GPU proof rejected by the restated verifier (code %d)" % rc)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        alg = algorithmic_counts()
        ntt_ms = micro["ifft"] + micro["lde"]
        hash_ms = micro["leaf_hash"] + micro["merkle_levels"]
        achieved = alg["ntt_bytes"] / (ntt_ms * 1e-3) / 1e9
        peak = float(peaks["hbm_gbs"])
        total = steps_total * world
        h2d = int(circ["wires"].nbytes + SALT_SEED.nbytes)   # e2e: witness + the 32-byte salt seed
        line = {
            "metric": METRIC, "value": total / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Goldilocks)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "degree_bits": PROOF_K, "zero_knowledge": PROOF_ZK,
                       "proofs_per_step": S * world, "step": "one proof on each of the %d streams of every GPU" % S,
                       "streams_per_gpu": S, "host_threads_per_gpu": HT, "host_wait": sync, "host_cores": cores,
                       "l2": "each proof streams ~0.3 GB of LDE/digest buffers through HBM (> 126 MB L2); "
                             "the commit microbench rotates 4 distinct 70.8 MB traces",
                       "parallelism": "independent proofs, %d stream(s) per GPU, no collective" % S,
                       "salts": "value: three [4][2^17] salt arrays resident in HBM; e2e: drawn on the device from a "
                                "32-byte seed (ChaCha8), as the reference draws its own from the OS RNG"},
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d * S, "d2h_bytes_per_step": len(proof) * S},
            "gpu_launches": int(launches),
            "single_proof_latency_ms": latency_ms,
            "proof_stage_ms": proof_stages,
            "circuit_constants_sigmas_commit_ms": circuit_commit_ms,
            "commit_microbench": {"workload": "PolynomialBatch::from_values 2^16 x 135, rate_bits=3, cap_height=4 "
                                              "(BASELINE configs[2])", "ms": sum(micro.values()), "stage_ms": micro,
                                  "blinding": {"workload": "the same with blinding = true: 4 salt columns (seed "
                                                           "0x5eed0002), 139-element leaves = 18 permutations",
                                               "ms": sum(micro_blind.values()), "stage_ms": micro_blind}},
            "roofline": {"bound": "hbm", "kernel": "commit microbench: IFFT (k_ntt_pass_a / k_ntt_pass_b_transpose) + "
                                                   "coset LDE (k_ntt_cluster)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of the three NTT launches of one commit: the
                         # LDE is one pass over HBM (cluster kernel, tile in distributed shared memory), the IFFT
                         # two; the kernels are integer-issue-bound, not HBM-bound
                         "traffic": NTT_DRAM_BYTES_PER_COMMIT,
                         "traffic_source": "profiles/r1_ntt_v2_ncu_full.txt (IFFT) + profiles/r2_ntt_cluster_2p16_ncu_full.txt (LDE)",
                         "peak_source": peak_src, "algorithmic_bytes": alg["ntt_bytes"],
                         "stage_ms": ntt_ms,
                         "note": "the north star's split: HBM roofline for the NTT/transpose stages (integer-issue-bound "
                                 "in practice, DESIGN.md 4), integer-multiplier roofline for Poseidon - the dominant "
                                 "kernel of the step - under roofline_int"},
            "roofline_int": {"bound": "int32-multiply", "kernel": "commit microbench: k_leaf_hash + k_merkle_level "
                                                                  "(Poseidon; the dominant kernel of every step)",
                             "achieved": alg["mults"] / (hash_ms * 1e-3) / 1e12, "peak": imad_wide / 1e12,
                             "unit": "T mul32/s", "frac": alg["mults"] / (hash_ms * 1e-3) / imad_wide,
                             "peak_source": "measured here: IMAD.WIDE.U32 (32x32+64, the instruction the field "
                                            "multiply and the MDS layer issue) on 8 independent accumulator chains "
                                            "per thread, SASS-checked; 6612 such multiplies per permutation is the "
                                            "algorithmic count of SURVEY 8(d) (its 2304 small-constant MDS "
                                            "multiplies now execute as DFMA on the FP64 pipe)",
                             "peak_imad_32bit": imad_lo / 1e12, "permutations": alg["perms"],
                             # the same work charged per instruction class: the 4308 wide (32x32+64) multiplies of a
                             # permutation against the IMAD.WIDE peak, its 2304 small-constant MDS multiplies against the
                             # 32-bit IMAD peak they were counted as in SURVEY 8(d) (they execute as DFMA today)
                             "frac_class_weighted": alg["perms"] * (4308.0 / imad_wide + 2304.0 / imad_lo) / (hash_ms * 1e-3),
                             "frac_wide_multiplies_only": alg["perms"] * 4308.0 / imad_wide / (hash_ms * 1e-3),
                             "ncu": {"source": "profiles/r2_poseidon_base_ncu_full.txt (the permutation of k_leaf_hash in the "
                                               "leaf-hash-shaped harness, 8.07 ms for the same 8.9 M permutations); "
                                               "SASS histogram: profiles/r2_leaf_hash_sass_histogram.txt",
                                     "issue_slots_busy_pct": 60.3, "fmaheavy_pipe_pct": 77.7, "alu_pipe_pct": 51.4,
                                     "fp64_pipe_pct": 11.7, "eligible_warps_per_cycle": 2.04,
                                     "warp_instructions_per_permutation": 19800,
                                     "limiter": "the FMA-heavy (integer multiply) pipe: 25.9 k pipe cycles per warp-permutation "
                                                "(IMAD.WIDE / IMAD.HI hold it 4 cycles, other IMAD forms 2) of 33.6 k elapsed"},
                             "traffic": LEAF_HASH_DRAM_BYTES_PER_COMMIT,
                             "perms_per_s": alg["perms"] / (hash_ms * 1e-3), "stage_ms": hash_ms},
            "clocks": clocks,
        }
        line["commit_microbench"]["ms_at_2^k_rows"] = sweep
        if voting is not None:
            line["voting_single_proof"] = voting
        if restored is not None:
            line["circuit_restored_from_files"] = restored
        if aggregator is not None:
            line["aggregator_node_proof"] = aggregator
        if cpu is not None:
            line["cpu_baseline"] = cpu
        # the multi-GPU results go LAST so that a tail of the line keeps them
        if tree is not None:
            line["aggregation_tree"] = tree
        if sharded is not None:
            line["commit_microbench_sharded"] = sharded
        if sharded_proof is not None:
            line["aggregation_proof_sharded"] = sharded_proof
        print(json.dumps(line), flush=True)

    for c, d, ds in zip(ctxs, dev_w, dev_s):
        c.dev_free(d)
        for p in ds:
            c.dev_free(p)
    for q in circuits:
        q.free()
    for c in ctxs:
        c.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16, help="a step = one proof on every stream of every GPU")
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--streams", type=int, default=8,
                    help="proofs in flight per GPU (measured at 128 steps: 6: 204, 8: 207, 12: 208 proofs/s; 8 divides the default step count)")
    ap.add_argument("--host-threads", type=int, default=1,
                    help="host threads per GPU driving the streams through qpzk_prove_begin / qpzk_prove_end")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sync", default="auto", choices=["auto", "spin", "yield", "blocking"],
                    help="how a proving thread waits for its stream: spin on a core, poll + sched_yield "
                         "(QPZK_CTX_YIELD_SYNC) or sleep on a blocking-sync event (QPZK_CTX_BLOCKING_SYNC); "
                         "auto = spin while streams x ranks fit the host cores, else yield")
    ap.add_argument("--no-aggregator", action="store_true", help="skip the aggregation proofs (configs[4])")
    ap.add_argument("--shard-k", type=int, default=18, help="log2 rows of the aggregation proof sharded over the GPUs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    if world == 1 and args.gpus > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--streams", str(args.streams), "--sync", args.sync, "--host-threads", str(args.host_threads),
               "--shard-k", str(args.shard_k)]
        cmd += ["--no-cpu"] * args.no_cpu + ["--no-aggregator"] * args.no_aggregator
        sys.exit(subprocess.call(cmd))
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
