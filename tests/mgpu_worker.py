"""Multi-rank worker: rows sharded over ranks, Phi^T-type results combined by allreduce.

backend gloo  -> CPU: each rank runs the ORACLE on its row block and the partial sums are combined
                 with torch.distributed (tests the sharding arithmetic and bench.py's row generator);
backend nccl  -> GPU: each rank runs the PRODUCT on its own B200, allreduce inside the C library
                 (ncclAllReduce on the library's stream); compared with the single-rank oracle.
Launched by tests/test_multirank.py (gloo, world 2) and by `torchrun` on a multi-GPU box.
"""
import os
import sys
from pathlib import Path

# the oracle's OpenMP regions ask for omp_get_num_procs() threads each: with one process per GPU they would
# oversubscribe the host world-size-fold and spin in each other's barriers (an 8-rank run did not finish in 150 s)
_world = int(os.environ.get("WORLD_SIZE", "1"))
os.environ.setdefault("OMP_THREAD_LIMIT", str(max(2, (os.cpu_count() or 2) // max(_world, 1))))
os.environ.setdefault("OMP_WAIT_POLICY", "passive")

import numpy as np
import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))
from conftest import make_problem, relerr  # noqa: E402
from outerbase_b200.binding import Library  # noqa: E402
import bench  # noqa: E402


def main():
    backend = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    oracle = Library(REPO / "oracle" / "_build" / "libob_oracle.so", "orc_")
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        import outerbase_b200 as obp
        lib = obp.lib(local)
        box = [lib.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        lib.comm_init(world, rank, box[0])
        assert lib.comm_info() == (world, rank)
    else:
        dist.init_process_group("gloo")
        lib = oracle

    # bench.py's row generator gives the same rows for any sharding
    full = bench.synth_rows(0, 5000, 4)
    lo, hi = (5000 * rank) // world, (5000 * (rank + 1)) // world
    np.testing.assert_array_equal(bench.synth_rows(lo, hi, 4), full[lo:hi])

    N, K = 6001, 300
    omr, x, y, terms, rng = make_problem(oracle, N, K, covs=["mat25"] * 8, knots=[np.arange(0.001, 0.999, 0.05)] * 8)
    lo, hi = (N * rank) // world, (N * (rank + 1)) // world
    om, *_ = make_problem(lib, N, K, covs=["mat25"] * 8, knots=[np.arange(0.001, 0.999, 0.05)] * 8)
    a, r = rng.normal(size=K) / 50, rng.normal(size=N)
    ref_ob = oracle.outerbase(omr, x)
    ob = lib.outerbase(om, x[lo:hi])
    # Phi a: row-local, no communication
    assert relerr(ob.matmul(terms, a), ref_ob.matmul(terms, a)[lo:hi]) < 1e-9
    # Phi^T r: partial sums over the rank's rows, summed over ranks
    part = ob.tmatmul(terms, r[lo:hi])
    if backend == "gloo":
        t = torch.from_numpy(part.copy()); dist.all_reduce(t); part = t.numpy()
    assert relerr(part, ref_ob.tmatmul(terms, r)) < 1e-9
    if backend == "nccl":
        # the cross-GPU sum itself: the one-shot peer-memory kernel (tagged 8-byte words, every rank adds the slots in
        # rank order) against torch's own NCCL allreduce, bit-identical across ranks; a payload above its capacity and
        # option p2p=0 both go through ncclAllReduce inside the library
        for n, p2p in ((K, 1), (1, 1), (4097, 1), (70000, 1), (K, 0)):
            lib.set_option("p2p", p2p)
            for rep in range(3):  # consecutive calls alternate the slot parity
                v = torch.from_numpy(np.random.default_rng([rank, n, rep]).normal(size=n)).cuda()
                want = v.clone(); dist.all_reduce(want)
                torch.cuda.synchronize()
                lib.allreduce_dev(v.data_ptr(), n); lib.synchronize()
                assert torch.allclose(v, want, rtol=1e-14, atol=1e-14), (n, p2p, rep)
                tmax, tmin = v.clone(), v.clone()
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
                assert torch.equal(tmax, tmin), (n, p2p, rep)
        lib.set_option("p2p", 1)
        # a rank WITHOUT rows takes the plain variant of the kernel while its peers run the fused reduce + allreduce
        nb = N if rank == 0 else 0
        obe = lib.outerbase(om, x[:nb] if rank == 0 else x[:0])
        part = obe.tmatmul(terms, r[:nb])
        assert relerr(part, ref_ob.tmatmul(terms, r)) < 1e-9

    # loglik_gauss / optcg on sharded rows
    ref = oracle.lpdfvec(oracle.logpr_gauss(omr, terms), oracle.loglik_gauss(omr, terms, y, x))
    ref.optcg(0.001, 100)
    if backend == "nccl":
        loglik = lib.loglik_gauss(om, terms, y[lo:hi], x[lo:hi])
        vec = lib.lpdfvec(lib.logpr_gauss(om, terms), loglik)
        # default noise scale log(0.01 var(y)) is computed over ALL ranks' rows inside the library
        assert abs(vec.para[1] - ref.para[1]) < 1e-12 * abs(ref.para[1]), (vec.para, ref.para)
        vec.updatepara(ref.para)
        vec.optcg(0.001, 100)
        assert vec.cg_iters == ref.cg_iters, (vec.cg_iters, ref.cg_iters)
        assert relerr(vec.coeff, ref.coeff) < 1e-8
        assert abs(vec.val - ref.val) <= 1e-8 * abs(ref.val)
        assert relerr(vec.gradhyp, ref.gradhyp) < 1e-6
        # every rank holds bit-identical results (allreduce gives all ranks the same bits)
        t = torch.from_numpy(vec.coeff.copy()).cuda()
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        assert torch.equal(tmax, tmin)
        # hyper-parameter learning on sharded rows (BASELINE config C4's driver): every rank runs the same BFGS_lpdf
        # control flow on bit-identical objective values; compared with the single-rank oracle run
        from outerbase_b200 import fitting
        ref.domarg = True; vec.domarg = True
        ro = fitting.BFGS_lpdf(omr, ref)
        rg = fitting.BFGS_lpdf(om, vec)
        assert abs(rg["optid"]["val"] - ro["optid"]["val"]) <= 1e-6 * abs(ro["optid"]["val"]), (rg["optid"]["val"], ro["optid"]["val"])
        assert relerr(rg["parlist"]["hyp"], ro["parlist"]["hyp"]) < 1e-4
        t = torch.tensor([rg["optid"]["val"]], dtype=torch.float64).cuda()
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        assert torch.equal(tmax, tmin)
    else:
        # the reduction loglik_gauss::update needs: grad (K) | ssq | row count, one allreduce
        lk = oracle.loglik_gauss(omr, terms, y[lo:hi], x[lo:hi])
        lk.updatepara(ref.para[1:2])
        lk.compute_gradpara = True
        lk.update(ref.coeff)
        sd = np.exp(ref.para[1])
        buf = torch.from_numpy(np.concatenate([lk.grad, [lk.gradpara[0] + (hi - lo)], [hi - lo]]))
        dist.all_reduce(buf)
        full_lk = oracle.loglik_gauss(omr, terms, y, x)
        full_lk.updatepara(ref.para[1:2]); full_lk.compute_gradpara = True; full_lk.update(ref.coeff)
        assert relerr(buf[:K].numpy(), full_lk.grad) < 1e-11
        ssq, n = float(buf[K]), float(buf[K + 1])
        assert n == N
        assert abs((-0.5 * ssq - n * np.log(sd)) - full_lk.val) <= 1e-11 * abs(full_lk.val)
    dist.barrier()
    if rank == 0:
        print(f"mgpu_worker[{backend}] world={world} OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
