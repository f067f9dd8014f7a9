"""Parity of the sharded path on real GPUs (run under torchrun, 2/4/8 ranks):
ShardedState.run == Circuit.run on one GPU, for circuits that need exchanges, local permutes and
rank-specialised gates.  Rank 0 prints 'dist_check ok' on success."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import qcircuits_b200 as qc
    from qcircuits_b200.dist import ShardedState
    from qcircuits_b200.workloads import layered_circuit

    rank = int(os.environ['RANK'])
    world = int(os.environ['WORLD_SIZE'])
    local_rank = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    g = int(np.log2(world))
    worst = 0.0
    for dtype, tol in ((np.complex64, 2e-5), (np.complex128, 1e-12)):
        for n in (14, 19, 22):
            c = layered_circuit(n, 6, seed=11 + n)
            c.append(qc.Toffoli(), [0, n - 1, 3])
            c.append(qc.Swap(), [0, n - 1])
            c.append(qc.ControlledU(np.exp(0.4j) * qc.RotationZ(0.8)), [1, 0])
            for q in range(n):
                c.append(qc.Hadamard(), [q])
            st = ShardedState(n, dtype)
            st.set_zeros().run(c)
            total = st.norm2()
            marg = st.marginals([0, n - 1, 2, 1])
            shard = qc._lib.download(st.engine.buf, dtype)
            shards = [None] * world
            dist.all_gather_object(shards, shard)
            if rank == 0:
                full = np.concatenate(shards)
                idx = np.arange(2 ** n)
                src = np.zeros_like(idx)
                for q in range(n):
                    src |= ((idx >> (n - 1 - q)) & 1) << st.phys[q]
                got = full[src] / np.sqrt(total)
                ref = c.run(qc.zeros(n, dtype=dtype))
                want = ref.to_column_vector()
                err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
                worst = max(worst, err / tol)
                assert err < tol, (n, dtype, err)
                pm = ref._measurement_probabilities([0, n - 1, 2, 1])
                assert np.max(np.abs(pm - marg)) < 10 * tol, (n, dtype)
                assert st.exchanges >= 1
            dist.barrier()
    if rank == 0:
        print('dist_check ok: world %d, worst error / tolerance = %.3f' % (world, worst))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
