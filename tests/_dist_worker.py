"""Worker for tests/test_dist_gloo.py (world_size 2, gloo, CPU only).

Exercises the host side of the multi-GPU path exactly as bench.py / the engine
drive it: gmm_shard_range -> per-shard packed statistics -> ONE all-reduce of the
packed buffer (here torch.distributed/gloo in place of ncclAllReduce) ->
replicated gmm_host_finalize on every rank.  The per-shard statistics come from
numpy (double) since no GPU is present; the oracle provides the responsibilities
and the single-process expected result (test infrastructure)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import entry, random_spd_params, assert_params_close  # noqa: E402
from test_host import numpy_stats  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = entry.load_package()
    orc = entry.load_oracle("f64")
    N, D, K = 2501, 5, 4            # odd N: the remainder goes to the last shard
    ev = pkg.synth.make_blobs(N, D, K, seed=17)
    ref = random_spd_params(pkg, K, D, np.random.default_rng(3))
    ref.memberships = np.zeros((K, N), np.float32)
    orc.constants(ref, K)
    soa = orc.transpose(ev)
    ll_ref = orc.estep(soa, ref, K)
    memb = ref.memberships.copy()
    got = ref.copy()
    orc.mstep(soa, ref, K)
    orc.constants(ref, K)

    b, n = pkg.shard_range(N, world, rank)
    shift = ev.astype(np.float64).mean(0)               # identical on all ranks (seeding all-reduce in the engine)
    stats = numpy_stats(pkg, ev[b:b + n], memb[:, b:b + n], shift, K)
    # local log-likelihood of the shard rides in the last slot of the same buffer
    cl_local = got.copy()
    cl_local.memberships = np.zeros((K, n), np.float32)
    stats[-1] = orc.estep(orc.transpose(np.ascontiguousarray(ev[b:b + n])), cl_local, K)
    assert stats.size == pkg.stats_len(K, D)
    t = torch.from_numpy(stats)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    pkg.host_finalize(t.numpy(), shift, got, K)
    assert_params_close(got, ref, K)
    assert abs(t.numpy()[-1] - ll_ref) < 2e-5 * abs(ll_ref), (t.numpy()[-1], ll_ref)
    # every rank must hold bit-identical parameters (replicated finalisation)
    for f in ("N", "means", "R", "Rinv", "constant", "pi"):
        a = torch.from_numpy(np.ascontiguousarray(getattr(got, f)).copy())
        lo, hi = a.clone(), a.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), f
    dist.barrier()
    if rank == 0:
        print("DIST_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
