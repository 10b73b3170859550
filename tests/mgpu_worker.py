"""Multi-rank worker: rows sharded over ranks, Phi^T-type results combined by allreduce.

backend gloo  -> CPU: each rank runs the ORACLE on its row block and the partial sums are combined
                 with torch.distributed (tests the sharding arithmetic and bench.py's row generator);
backend nccl  -> GPU: each rank runs the PRODUCT on its own B200, allreduce inside the C library
                 (ncclAllReduce on the library's stream); compared with the single-rank oracle.
Launched by tests/test_multirank.py (gloo, world 2) and by `torchrun` on a multi-GPU box.

The single-rank ORACLE reference (full-data products, optcg, a whole BFGS_lpdf run) is computed by rank 0 alone, with all
host threads, BEFORE the process group exists, and handed to the other ranks through a file: with one oracle run per
rank an 8-rank launch oversubscribed the host eight-fold and did not finish.
"""
import os
import sys
from pathlib import Path

# the oracle's OpenMP regions ask for omp_get_num_procs() threads each: with one process per GPU they would
# oversubscribe the host world-size-fold and spin in each other's barriers (an 8-rank run did not finish in 150 s)
_world = int(os.environ.get("WORLD_SIZE", "1"))
if os.environ.get("RANK", "0") == "0":  # rank 0 computes the oracle reference alone while its peers sleep: whole machine
    os.environ.pop("OMP_THREAD_LIMIT", None)
    os.environ.pop("OMP_NUM_THREADS", None)
else:
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


PROBLEM = dict(covs=["mat25"] * 8, knots=[np.arange(0.001, 0.999, 0.05)] * 8)
N, K = 6001, 300


def oracle_reference(oracle, with_bfgs):
    """Everything the single-rank oracle has to say about the test problem (rank 0 only)."""
    from outerbase_b200 import fitting
    omr, x, y, terms, rng = make_problem(oracle, N, K, **PROBLEM)
    a, r = rng.normal(size=K) / 50, rng.normal(size=N)
    ref_ob = oracle.outerbase(omr, x)
    out = dict(matmul=ref_ob.matmul(terms, a), tmatmul=ref_ob.tmatmul(terms, r))
    ref = oracle.lpdfvec(oracle.logpr_gauss(omr, terms), oracle.loglik_gauss(omr, terms, y, x))
    ref.optcg(0.001, 100)
    out.update(para=np.array(ref.para), cg_iters=ref.cg_iters, coeff=np.array(ref.coeff), val=ref.val, gradhyp=np.array(ref.gradhyp))
    full_lk = oracle.loglik_gauss(omr, terms, y, x)
    full_lk.updatepara(ref.para[1:2]); full_lk.compute_gradpara = True; full_lk.update(ref.coeff)
    out.update(lk_grad=np.array(full_lk.grad), lk_val=full_lk.val)
    if with_bfgs:
        ref.domarg = True
        ro = fitting.BFGS_lpdf(omr, ref)
        out.update(bfgs_val=ro["optid"]["val"], bfgs_hyp=np.array(ro["parlist"]["hyp"]))
    return out


def shared_reference(oracle, rank, with_bfgs):
    import time
    path = Path(os.environ.get("OB_MGPU_REF", f"/tmp/ob_mgpu_ref_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}.npz"))
    if rank == 0:
        ref = oracle_reference(oracle, with_bfgs)
        tmp = path.with_suffix(".tmp.npz")
        np.savez(tmp, **ref)
        os.replace(tmp, path)
        return ref
    t0 = time.time()
    while not path.exists():
        if time.time() - t0 > 500:
            raise TimeoutError(f"rank 0 did not deliver {path}")
        time.sleep(0.25)
    return dict(np.load(path))


def main():
    backend = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    oracle = Library(REPO / "oracle" / "_build" / "libob_oracle.so", "orc_")
    R = shared_reference(oracle, rank, with_bfgs=(backend == "nccl"))
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

    lo, hi = (N * rank) // world, (N * (rank + 1)) // world
    om, x, y, terms, rng = make_problem(lib, N, K, **PROBLEM)  # selectterms is bit-exact across the two libraries (test_host_model)
    a, r = rng.normal(size=K) / 50, rng.normal(size=N)
    ob = lib.outerbase(om, x[lo:hi])
    # Phi a: row-local, no communication
    assert relerr(ob.matmul(terms, a), R["matmul"][lo:hi]) < 1e-9
    # Phi^T r: partial sums over the rank's rows, summed over ranks
    part = ob.tmatmul(terms, r[lo:hi])
    if backend == "gloo":
        t = torch.from_numpy(part.copy()); dist.all_reduce(t); part = t.numpy()
    assert relerr(part, R["tmatmul"]) < 1e-9
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
        assert relerr(part, R["tmatmul"]) < 1e-9

    # loglik_gauss / optcg on sharded rows
    if backend == "nccl":
        loglik = lib.loglik_gauss(om, terms, y[lo:hi], x[lo:hi])
        vec = lib.lpdfvec(lib.logpr_gauss(om, terms), loglik)
        # default noise scale log(0.01 var(y)) is computed over ALL ranks' rows inside the library
        assert abs(vec.para[1] - R["para"][1]) < 1e-12 * abs(R["para"][1]), (vec.para, R["para"])
        vec.updatepara(R["para"])
        vec.optcg(0.001, 100)
        assert vec.cg_iters == int(R["cg_iters"]), (vec.cg_iters, R["cg_iters"])
        assert relerr(vec.coeff, R["coeff"]) < 1e-8
        assert abs(vec.val - float(R["val"])) <= 1e-8 * abs(float(R["val"]))
        assert relerr(vec.gradhyp, R["gradhyp"]) < 1e-6
        # every rank holds bit-identical results (allreduce gives all ranks the same bits)
        t = torch.from_numpy(vec.coeff.copy()).cuda()
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        assert torch.equal(tmax, tmin)
        # hyper-parameter learning on sharded rows (BASELINE config C4's driver): every rank runs the same BFGS_lpdf
        # control flow on bit-identical objective values; compared with the single-rank oracle run
        from outerbase_b200 import fitting
        vec.domarg = True
        rg = fitting.BFGS_lpdf(om, vec)
        assert abs(rg["optid"]["val"] - float(R["bfgs_val"])) <= 1e-6 * abs(float(R["bfgs_val"])), (rg["optid"]["val"], R["bfgs_val"])
        assert relerr(rg["parlist"]["hyp"], R["bfgs_hyp"]) < 1e-4
        t = torch.tensor([rg["optid"]["val"]], dtype=torch.float64).cuda()
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        assert torch.equal(tmax, tmin)
    else:
        # the reduction loglik_gauss::update needs: grad (K) | ssq | row count, one allreduce
        lk = oracle.loglik_gauss(om, terms, y[lo:hi], x[lo:hi])
        lk.updatepara(R["para"][1:2])
        lk.compute_gradpara = True
        lk.update(R["coeff"])
        sd = np.exp(R["para"][1])
        buf = torch.from_numpy(np.concatenate([lk.grad, [lk.gradpara[0] + (hi - lo)], [hi - lo]]))
        dist.all_reduce(buf)
        assert relerr(buf[:K].numpy(), R["lk_grad"]) < 1e-11
        ssq, n = float(buf[K]), float(buf[K + 1])
        assert n == N
        assert abs((-0.5 * ssq - n * np.log(sd)) - float(R["lk_val"])) <= 1e-11 * abs(float(R["lk_val"]))
    dist.barrier()
    if rank == 0:
        print(f"mgpu_worker[{backend}] world={world} OK")
        for f in Path("/tmp").glob(f"ob_mgpu_ref_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}*.npz"):
            f.unlink(missing_ok=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
