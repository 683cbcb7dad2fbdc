# torchrun script: latency of ONE 2^k-row recursion-shaped node proof spread over the ranks (host buffers, pinned),
# against the same proof on one GPU. Is sharding the upper levels of the aggregation tree worth it?
import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/qp-zk-circuits-rm_b200")
import numpy as np, torch, torch.distributed as dist
import qpzk
from qpzk import synth, dist as qdist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = qpzk.Context(lr)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 13
bc = synth.build_recursion(k, zk=True, seed=9, provider=synth.GpuProvider(ctx))
circ = qpzk.Circuit(ctx, bc["common"], bc["digest"], bc["constants_sigmas"])
pw = qpzk.PinnedBuffer(bc["wires"].shape); pw.array[...] = bc["wires"]
ps = [qpzk.PinnedBuffer(s.shape) for s in bc["salts"]]
for a, s in zip(ps, bc["salts"]): a.array[...] = s
salts = [a.array for a in ps]
groups = {}
g = world
while g >= 1:
    # every rank creates every group (new_group is collective); ranks [j*g, (j+1)*g) form one
    for j in range(world // g):
        grp = dist.new_group(list(range(j * g, (j + 1) * g)))
        if j * g <= rank < (j + 1) * g:
            groups[g] = grp
    g //= 2
alone = circ.prove(pw.array, bc["public_inputs"], salts)
for g in sorted(groups):
    times = []
    for i in range(8):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if g == 1:
            p = circ.prove(pw.array, bc["public_inputs"], salts)
        else:
            p = qdist.prove_sharded_nccl(circ, pw.array, bc["public_inputs"], salts, 4, 3, on_device=False, group=groups[g])
        ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if i >= 2: times.append(float(ms.item()))
        assert p == alone
    if rank == 0:
        print("2^%d-row node over %d GPU(s): median %.2f ms min %.2f  stages %s" % (k, g, float(np.median(times)), min(times), {a: round(b, 2) for a, b in circ.stage_ms().items()}), flush=True)
dist.barrier()
dist.destroy_process_group()
