"""bench.py body for N > 1 (launched by torchrun, one rank per GPU): the sharded layered circuit.

Weak scaling: every rank holds 2^33 complex64 amplitudes (64 GiB) -- 34 qubits on 2 GPUs, 35 on 4,
36 on 8 (BASELINE.json: "36q at 8 GPUs"; 36 qubits do not fit 2 or 4 B200s).
"""
import json
import os
import time

import numpy as np

SEED = 1234


def main(args, parity_check=None):
    import torch
    import torch.distributed as dist
    from . import _lib
    from .dist import ShardedState, dist_schedule, _LogicalGate
    from .workloads import layered_circuit
    from .benchutil import ClockSampler, measured_peak

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', str(rank)))
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    if parity_check is not None:
        args.parity = parity_check()
    n, depth = args.qubits, args.depth
    g = int(np.log2(world))
    np_dtype = np.complex64 if args.dtype == 'c64' else np.complex128
    amp_bytes = 8 if args.dtype == 'c64' else 16
    circ = layered_circuit(n, depth, seed=SEED)
    ngates = len(circ)
    st = ShardedState(n, np_dtype)
    shard_bytes = (2 ** st.n_local) * amp_bytes

    def step():
        st.set_zeros()
        st.run(circ, max_ops=args.max_ops)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    total = st.norm2()
    launches0, exch0 = st.engine.launches, st.exchanges
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    passes = (st.engine.launches - launches0) / args.steps
    exchanges = (st.exchanges - exch0) / args.steps
    marg = st.marginals([0, n // 2, n - 1])

    # exchange alone, timed on the device (max over ranks)
    torch.cuda.synchronize()
    dist.barrier()
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x0.record()
    st.engine.exchange(None)
    st.engine.exchange(None)
    x1.record()
    torch.cuda.synchronize()
    xms = torch.tensor([x0.elapsed_time(x1) / 2], dtype=torch.float64, device='cuda')
    dist.all_reduce(xms, op=dist.ReduceOp.MAX)
    exchange_ms = float(xms.item())
    exchange_kind = getattr(st.engine, 'exchange_kind', None)

    # end to end: every rank's shard comes from host memory and the result goes back to host memory inside the timed
    # region, through a BOUNDED pinned staging ring (4 x 256 MiB per rank) -- whole pinned shards (2 x 64 GiB per rank)
    # do not fit a host.  The input is the workload's |0...0> (rank 0's first chunk holds the 1); every byte of every
    # shard crosses PCIe in both directions.
    e2e = None
    if not args.skip_e2e:
        from .dist import initial_layout
        ring_chunks, chunk_elems = 4, (256 << 20) // st.engine.buf.element_size()
        chunk_elems = min(chunk_elems, st.engine.buf.numel())
        ring = [torch.zeros(chunk_elems, dtype=st.engine.buf.dtype).pin_memory() for _ in range(ring_chunks)]
        first = torch.zeros(chunk_elems, dtype=st.engine.buf.dtype).pin_memory()
        if rank == 0:
            first[0] = 1.0
        copy_stream = torch.cuda.Stream()
        nchunks = st.engine.buf.numel() // chunk_elems
        checks = torch.zeros(1, dtype=torch.float64)

        def stream_in():
            buf = st.engine.buf
            for c in range(nchunks):
                src_chunk = first if c == 0 else ring[c % ring_chunks]
                buf[c * chunk_elems:(c + 1) * chunk_elems].copy_(src_chunk, non_blocking=True)

        def stream_out():
            buf = st.engine.buf
            evs = [None] * ring_chunks
            for c in range(nchunks):
                slot = c % ring_chunks
                if evs[slot] is not None:
                    evs[slot].synchronize()          # the host "consumes" the slot before it is reused
                ring[slot].copy_(buf[c * chunk_elems:(c + 1) * chunk_elems], non_blocking=True)
                evs[slot] = torch.cuda.Event()
                evs[slot].record()
            torch.cuda.current_stream().synchronize()
            ring[0].zero_(); ring[1].zero_(); ring[2].zero_(); ring[3].zero_()       # back to the |0..0> input image

        def e2e_step():
            st.phys = list(initial_layout(n, g))
            stream_in()
            st.run(circ, max_ops=args.max_ops)
            stream_out()

        e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        reps = max(1, args.steps // 2)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            e2e_step()
        t1.record()
        torch.cuda.synchronize()
        tms = torch.tensor([t0.elapsed_time(t1) / reps], dtype=torch.float64, device='cuda')
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e = {'value': ngates / (float(tms.item()) * 1e-3), 'unit': 'gates/s', 'ms_per_step': float(tms.item()),
               'h2d_bytes_per_step': int(shard_bytes * world), 'd2h_bytes_per_step': int(shard_bytes * world),
               'staging': '%d x %d MiB pinned ring per rank' % (ring_chunks, chunk_elems * st.engine.buf.element_size() >> 20)}
        del ring, first

    # strong scaling: the SAME 33-qubit circuit on 1 / 2 / 4 / 8 GPUs (64 GiB in total; bench.py --gpus 1 reports the
    # single-GPU leg), device-resident like `value`
    strong = None
    if not args.skip_strong:
        if hasattr(st.engine, 'release_peers'):
            st.engine.release_peers()          # peers unmap this rank's shards before they are freed
        st.engine.buf = st.engine.spare = None          # 2 x 64 GiB back before the next state is allocated
        del st
        torch.cuda.empty_cache()
        ns = args.strong_qubits
        circ_s = layered_circuit(ns, depth, seed=SEED)
        st_s = ShardedState(ns, np_dtype)
        for _ in range(2):
            st_s.set_zeros()
            st_s.run(circ_s, max_ops=args.max_ops)
        torch.cuda.synchronize()
        dist.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            st_s.set_zeros()
            st_s.run(circ_s, max_ops=args.max_ops)
        s1.record()
        torch.cuda.synchronize()
        sms = torch.tensor([s0.elapsed_time(s1) / args.steps], dtype=torch.float64, device='cuda')
        dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        strong = {'qubits': ns, 'gates': len(circ_s), 'ms_per_step': float(sms.item()),
                  'gates_per_s': len(circ_s) / (float(sms.item()) * 1e-3), 'norm_check': st_s.norm2(),
                  'amp_updates_per_s_per_gpu': len(circ_s) * 2.0 ** ns / (float(sms.item()) * 1e-3) / world}
        if hasattr(st_s.engine, 'release_peers'):
            st_s.engine.release_peers()
        del st_s


    if rank == 0:
        peak, peak_src = measured_peak()
        nvlink = 770.0   # GB/s per direction per GPU, measured peer copy (B200_PROFILING.md)
        pass_bytes = 2.0 * shard_bytes
        sent = shard_bytes * (1.0 - 1.0 / world)
        t_roof = (passes * pass_bytes / (peak * 1e9) + exchanges * sent / (nvlink * 1e9)) * 1e3
        achieved = pass_bytes * passes / (ms_per_step * 1e-3) / 1e9
        line = {
            'metric': 'gates/s, random layered circuit', 'value': ngates / (ms_per_step * 1e-3), 'unit': 'gates/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype,
            'data': 'synthetic',
            'config': {'workload': 'layered H/RX/RZ+CNOT/CZ circuit, %d qubits, depth %d, seed %d, |0..0> input, sharded '
                                   'on the top %d bits' % (n, depth, SEED, g),
                       'gates': ngates, 'shard_gib': shard_bytes / 2 ** 30, 'passes_per_step_rank0': passes,
                       'exchanges_per_step': exchanges, 'max_ops_per_pass': args.max_ops,
                       'l2': 'shard (%.0f GiB) is far larger than L2; no flush needed' % (shard_bytes / 2 ** 30)},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': None, 'peak_source': peak_src, 'kernel': 'qcs::tile_kernel_tc / tile_kernel_pipe (fused gate passes), per GPU',
                         'bytes_per_launch': pass_bytes,
                         'combined': {'hbm_plus_nvlink_roofline_ms': t_roof, 'measured_ms': ms_per_step,
                                      'frac': t_roof / ms_per_step, 'nvlink_gbs_per_dir': nvlink,
                                      'exchange_ms_measured': exchange_ms,
                                      'exchange_gbs_per_gpu': sent / (exchange_ms * 1e-3) / 1e9,
                                      'exchange_kernel': {'p2p': 'qcs::dist_pull_kernel (libqcs: P2P stores into the peers\' shards over '
                                                                 'NVLink, CUDA IPC; two one-element NCCL all-reduces order the ranks)',
                                                          'nccl': 'torch.distributed.all_to_all_single (NCCL)'}.get(
                                          exchange_kind, 'unknown')}},
            'cpu_baseline': None,
            # per-GPU work rate, comparable across N and with the single-GPU line: gates x 2^n amplitudes / time / GPUs
            'amp_updates_per_s_per_gpu': ngates * 2.0 ** n / (ms_per_step * 1e-3) / world,
            'strong_scaling': strong, 'parity': getattr(args, 'parity', None),
            'e2e': e2e, 'clocks': clocks.summary(),
            'gpu_launches': int(passes * args.steps),
            'norm_check': total, 'marginals_q0_mid_last': [float(x) for x in marg],
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
