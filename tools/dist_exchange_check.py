"""Multi-GPU: the qubit exchange of libqcs (qcs_dist.cu) checked and timed.

  torchrun --nproc-per-node N tools/dist_exchange_check.py          one process per GPU: CudaEngine.exchange through
        qcs_dist_swap_rank + CUDA IPC against torch.distributed.all_to_all_single (bit-exact), then both timed
  python tools/dist_exchange_check.py single                         one process, all visible GPUs: qcs_dist_init /
        qcs_dist_swap / qcs_dist_destroy against a NumPy index map (general bit pairs), then timed
Prints one JSON line per mode (rank 0)."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def per_rank():
    import torch
    import torch.distributed as dist
    from qcircuits_b200 import _lib
    from qcircuits_b200.dist import CudaEngine
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    dist.init_process_group('nccl')
    out = {'mode': 'one process per GPU', 'world': world}
    for dtype in (np.complex64, np.complex128):
        n_local = 16
        eng = CudaEngine(n_local, dtype)
        gen = torch.Generator(device='cuda').manual_seed(100 + rank)
        eng.buf.copy_(torch.rand(eng.buf.numel(), generator=gen, device='cuda', dtype=torch.float64).to(eng.buf.dtype))
        want = torch.empty_like(eng.buf)
        dist.all_to_all_single(want, eng.buf.clone())
        for rep in range(3):                    # three in a row: both buffers serve as source and destination
            before = eng.buf.clone()
            eng.exchange(None)
            assert eng.exchange_kind == 'p2p', eng.exchange_kind
            ref = torch.empty_like(before)
            dist.all_to_all_single(ref, before)
            assert torch.equal(eng.buf, ref), (str(dtype), rep)
        assert torch.equal(eng.buf, want)       # an odd number of exchanges = one exchange
        del eng
    # timing at a large shard (2^30 complex64 = 8 GiB per GPU unless QX_BITS says otherwise)
    n_local = int(os.environ.get('QX_BITS', '30'))
    times = {}
    for mode in ('p2p', 'nccl'):
        os.environ['QCS_DIST_EXCHANGE'] = mode
        eng = CudaEngine(n_local, np.complex64)
        eng.buf.zero_()
        for _ in range(2):
            eng.exchange(None)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 4
        for _ in range(reps):
            eng.exchange(None)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device='cuda')
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sent = (8 << n_local) * (world - 1) / world           # bytes that leave (and enter) each GPU
        times[mode] = {'kind': eng.exchange_kind, 'ms': float(ms.item()), 'GBps_per_gpu_each_way': sent / float(ms.item()) / 1e6}
        del eng
        torch.cuda.empty_cache()
    out['n_local'] = n_local
    out['exchange'] = times
    if rank == 0:
        print(json.dumps(out), flush=True)
        print('dist_exchange_check ok', flush=True)
    dist.destroy_process_group()


def single():
    import torch
    from qcircuits_b200 import _lib
    lib = _lib.load()
    ndev = torch.cuda.device_count()
    ndev = 1 << (ndev.bit_length() - 1)
    if ndev < 2:
        print('needs 2 GPUs')
        return
    G = ndev.bit_length() - 1
    devs = (ctypes.c_int * ndev)(*range(ndev))
    h = ctypes.c_void_p()
    _lib.check(lib.qcs_dist_init(ndev, devs, ctypes.byref(h)))
    streams = [torch.cuda.Stream(device=d) for d in range(ndev)]
    sarr = (ctypes.c_void_p * ndev)(*[s.cuda_stream for s in streams])
    rng = np.random.default_rng(3)
    out = {'mode': 'one process, all GPUs', 'ndev': ndev}
    for np_dtype, code, tdt in ((np.complex64, 0, torch.float32), (np.complex128, 1, torch.float64)):
        n_local = 14
        for trial in range(4):
            g = int(rng.integers(1, G + 1))
            gb = [int(x) for x in rng.permutation(G)[:g]]
            lo = 1 if code == 0 else 0
            lb = [int(x) for x in lo + rng.permutation(n_local - lo)[:g]]
            host = [(rng.normal(size=1 << n_local) + 1j * rng.normal(size=1 << n_local)).astype(np_dtype) for _ in range(ndev)]
            src, dst = [], []
            for d in range(ndev):
                with torch.cuda.device(d):
                    src.append(torch.from_numpy(host[d].view(np.float32 if code == 0 else np.float64)).to('cuda:%d' % d))
                    dst.append(torch.empty_like(src[-1]))
            for d in range(ndev):
                torch.cuda.synchronize(d)
            sp = (ctypes.c_void_p * ndev)(*[x.data_ptr() for x in src])
            dp = (ctypes.c_void_p * ndev)(*[x.data_ptr() for x in dst])
            _lib.check(lib.qcs_dist_swap(h, dp, sp, n_local, code, _lib.int32_array(gb), _lib.int32_array(lb), g, sarr))
            for d in range(ndev):
                torch.cuda.synchronize(d)
            i = np.arange(1 << n_local)
            for r in range(ndev):
                si, sr = i.copy(), np.full_like(i, r)
                for j in range(g):
                    diff = ((i >> lb[j]) ^ (r >> gb[j])) & 1
                    si ^= diff << lb[j]
                    sr ^= diff << gb[j]
                want = np.empty(1 << n_local, dtype=np_dtype)
                for p in range(ndev):
                    sel = sr == p
                    want[sel] = host[p][si[sel]]
                got = dst[r].cpu().numpy().view(np_dtype)
                assert np.array_equal(got, want), (np_dtype, trial, r, gb, lb)
    # timing: all rank bits <-> the top local bits, 2^30 complex64 per GPU
    n_local = int(os.environ.get('QX_BITS', '30'))
    src = [torch.zeros(2 << n_local, dtype=torch.float32, device='cuda:%d' % d) for d in range(ndev)]
    dst = [torch.empty_like(x) for x in src]
    sp = (ctypes.c_void_p * ndev)(*[x.data_ptr() for x in src])
    dp = (ctypes.c_void_p * ndev)(*[x.data_ptr() for x in dst])
    gb, lb = list(range(G)), [n_local - G + j for j in range(G)]
    ev = []
    for rep in range(5):
        if rep == 1:
            for d in range(ndev):
                torch.cuda.synchronize(d)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ndev)]
            for d in range(ndev):
                with torch.cuda.device(d):
                    ev[d][0].record(streams[d])
        a, b = (dp, sp) if rep % 2 == 0 else (sp, dp)
        _lib.check(lib.qcs_dist_swap(h, a, b, n_local, 0, _lib.int32_array(gb), _lib.int32_array(lb), G, sarr))
    for d in range(ndev):
        with torch.cuda.device(d):
            ev[d][1].record(streams[d])
    for d in range(ndev):
        torch.cuda.synchronize(d)
    ms = max(e0.elapsed_time(e1) for e0, e1 in ev) / 4
    sent = (8 << n_local) * (ndev - 1) / ndev
    out['n_local'] = n_local
    out['exchange'] = {'ms': ms, 'GBps_per_gpu_each_way': sent / ms / 1e6}
    _lib.check(lib.qcs_dist_destroy(h))
    print(json.dumps(out), flush=True)
    print('dist_exchange_single ok', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'single':
        single()
    else:
        per_rank()
