"""Parity of the sharded path on real GPUs (run under torchrun, 2/4/8 ranks) against the ORACLE:
ShardedState.run for circuits that need exchanges, local permutes and rank-specialised gates, and the drop-in
surface -- from_column_vector, Operator.__call__, marginal probabilities, measure (keep / remove, rank-bit and local
qubits, identical outcomes for identical uniforms), tensor product.  Rank 0 prints 'dist_check ok' and the achieved
error relative to the tolerance (1e-5 complex64, 1e-12 complex128)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def check(max_n=22, verbose=True):
    """runs inside an initialised process group; returns worst error / tolerance over all checks"""
    import torch.distributed as dist
    import qcircuits_b200 as qc
    from oracle import qc_oracle as orc
    from qcircuits_b200.dist import ShardedState
    from qcircuits_b200.workloads import layered_circuit

    rank = dist.get_rank()
    worst = 0.0
    for dtype, tol in ((np.complex64, 1e-5), (np.complex128, 1e-12)):
        for n in [s for s in (14, 19, 22) if s <= max_n]:
            c = layered_circuit(n, 6, seed=11 + n)
            c.append(qc.Toffoli(), [0, n - 1, 3])
            c.append(qc.Swap(), [0, n - 1])
            c.append(qc.ControlledU(np.exp(0.4j) * qc.RotationZ(0.8)), [1, 0])
            for q in range(n):
                c.append(qc.Hadamard(), [q])
            st = qc.zeros(n, dtype=dtype, sharded=True)
            assert isinstance(st, ShardedState) and st.rank == n
            st.run(c)
            want = np.zeros(2 ** n, dtype=np.complex128)
            want[0] = 1.0
            for g in c.gates:
                want = orc.apply_matrix_flat(g.plan.full, want, g.bits, renorm=False)
            want /= np.linalg.norm(want)
            got = st.to_column_vector()
            err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
            worst = max(worst, err / tol)
            assert err < tol, ('circuit', n, dtype, err)
            qs = [0, n - 1, 2, 1]
            ps, _, _ = orc.measurement_probabilities(want.reshape([2] * n), qs)
            perr = float(np.max(np.abs(st.marginal_probabilities(qs) - ps)))
            worst = max(worst, perr / tol)
            assert perr < tol, ('marginals', n, dtype, perr)
            assert st.exchanges >= 1
            # Operator.__call__ on a sharded state: new object, argument untouched
            rng = np.random.default_rng(n)
            A = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
            Q, R = np.linalg.qr(A)
            U = Q * (np.diag(R) / np.abs(np.diag(R)))
            st2 = qc.Operator.from_matrix(U)(st, qubit_indices=[0, n - 3])
            want2 = orc.apply_matrix_flat(U, want, [n - 1, 2], renorm=True)
            err = float(np.max(np.abs(st2.to_column_vector() - want2)) / np.max(np.abs(want2)))
            worst = max(worst, err / tol)
            assert err < tol, ('operator call', n, dtype, err)
            assert np.max(np.abs(st.to_column_vector() - want)) < tol
            # measurement: the same uniform -> the oracle's outcome and post-measurement state
            for mq, remove, u in (([0, 3], False, 0.37), ([n - 1], False, 0.81), ([0], True, 0.05), ([2, 0, n - 1], True, 0.66)):
                s3 = st2.copy()
                bits = s3.measure(mq, remove=remove, u=u)
                bits_want, post_want = orc.measure(want2.reshape([2] * n), mq, u, remove=remove)
                assert tuple(bits) == tuple(bits_want), ('measure outcome', n, dtype, mq, remove)
                post = post_want.reshape(-1)
                err = float(np.max(np.abs(s3.to_column_vector() - post)) / np.max(np.abs(post)))
                worst = max(worst, err / tol)
                assert err < 2 * tol, ('post-measurement state', n, dtype, mq, remove, err)   # (one more rounding: /sqrt(p))
            # from_column_vector and the tensor product with an ordinary state
            vec = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
            vec /= np.linalg.norm(vec)
            s4 = ShardedState.from_column_vector(vec, dtype)
            small = qc.State.from_column_vector(np.array([0.6, 0.8j, 0.0, 0.0]), dtype=dtype)
            for prod, wantp in ((s4 * small, np.kron(vec, small.to_column_vector())), (small * s4, np.kron(small.to_column_vector(), vec))):
                err = float(np.max(np.abs(prod.to_column_vector() - wantp)) / np.max(np.abs(wantp)))
                worst = max(worst, err / tol)
                assert err < tol, ('tensor product', n, dtype, err)
            dist.barrier()
    if rank == 0 and verbose:
        print('dist_check ok: world %d, worst error / tolerance = %.3f' % (dist.get_world_size(), worst))
    return worst


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ['RANK'])
    local_rank = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    check()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
