#!/usr/bin/env python
"""bench.py -- gates/s and HBM roofline of the state-vector hot path on B200.

Workload (BASELINE.json metric): seeded random layered circuit -- per layer one of
H / RX(theta) / RZ(theta) on every qubit, then a brick of CNOT or CZ on neighbouring qubits -- on
|0...0>, complex64, depth 40: 30 qubits on one GPU (1 780 gates).  A "step" is one execution of the
whole circuit.  Lines printed (one JSON object on stdout, rank 0 only):

  value      gates/s with the state resident in HBM (fused passes through Circuit.run)
  e2e        gates/s through the host-buffer API (Circuit.run_host): pinned host state -> device,
             circuit, normalised state -> pinned host, all inside the timed region
  roofline   achieved HBM GB/s of the dominant kernel (tile pass): algorithmic bytes of one pass
             (2 * 2^n * sizeof(amp)) / mean launch duration, against MEASURED_PEAKS.json
  cpu_baseline   the NumPy port of the reference algorithm (oracle/, kind "port") on the host cores,
             bounded sample, scaled to the workload size

`--impl reference` times only that CPU port (the reference is pure Python/NumPy and cannot travel
to the GPU box, so the oracle port stands in for it: same tensordot/transposes/renormalise passes).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 1234


from qcircuits_b200.benchutil import ClockSampler, measured_peak, recorded_traffic  # noqa: E402


def cpu_port_gates_per_s(orc, n_sample, n_target, budget_s, depth):
    """Time the oracle's reference-route gate application (tensordot + copy + renormalise) on the
    host: the first gates of the same circuit at n_sample qubits, scaled by 2^(n_sample - n_target)."""
    gates = orc.layered_circuit(n_sample, depth, seed=SEED)
    psi = orc.zeros_state(n_sample)
    psi = orc.apply_operator(orc.gate_tensor(*gates[0][:2]), psi, list(gates[0][2]))   # first touch, untimed
    done, t0 = 0, time.perf_counter()
    for name, theta, qs in gates[1:]:
        psi = orc.apply_operator(orc.gate_tensor(name, theta), psi, list(qs))
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    rate_sample = done / dt
    return rate_sample * 2.0 ** (n_sample - n_target), rate_sample, done, dt


def reference_arm(args):
    from oracle import qc_oracle as orc
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = args.qubits
    n_sample = min(args.cpu_qubits, n)
    per_step = []
    for _ in range(args.warmup + args.steps):
        scaled, raw, done, dt = cpu_port_gates_per_s(orc, n_sample, n, args.cpu_budget, args.depth)
        per_step.append((scaled, dt))
    timed = per_step[args.warmup:]
    value = float(np.mean([v for v, _ in timed]))
    sample = ('first gates of the same layered circuit at %d qubits complex128 (the reference has no complex64), '
              '%.0f s per step, gates/s scaled by 2^(%d-%d) to the %d-qubit workload' % (n_sample, args.cpu_budget, n_sample, n, n))
    line = {'impl': 'reference', 'metric': 'gates/s, random layered circuit', 'value': value, 'unit': 'gates/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * float(np.mean([d for _, d in timed])), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'c128', 'data': 'synthetic',
            'config': {'workload': 'layered H/RX/RZ+CNOT/CZ circuit, %d qubits, depth %d, seed %d, |0..0> input' % (n, args.depth, SEED)},
            'cpu_baseline': {'value': value, 'unit': 'gates/s', 'cores': 1, 'kind': 'port', 'sample': sample,
                             'host_cores': os.cpu_count()},
            'e2e': {'value': value, 'unit': 'gates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--qubits', type=int, default=None)
    ap.add_argument('--depth', type=int, default=40)
    ap.add_argument('--dtype', default='c64', choices=['c64', 'c128'])
    ap.add_argument('--max-ops', type=int, default=128)
    ap.add_argument('--no-fuse', action='store_true')
    ap.add_argument('--cpu-qubits', type=int, default=24)
    ap.add_argument('--cpu-budget', type=float, default=12.0)
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--skip-e2e', action='store_true')
    ap.add_argument('--workload', default='layered', choices=['layered', 'qpe'],
                    help='layered = the headline random circuit; qpe = BASELINE config 3 (phase estimation + inverse QFT)')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.qubits is None:
        args.qubits = 30 if world == 1 else 33 + int(np.log2(world))
    if args.workload == 'qpe':          # config 3 is a device-resident run; the CPU / host-buffer legs belong to the headline
        args.skip_cpu = args.skip_e2e = True
        if args.qubits is None or args.qubits == 30:
            args.qubits = 32
    if args.impl == 'reference':
        return reference_arm(args)
    if world > 1:
        from qcircuits_b200 import dist_bench
        return dist_bench.main(args)

    import torch
    import qcircuits_b200 as qc
    from qcircuits_b200.workloads import layered_circuit

    torch.cuda.set_device(0)
    n, depth = args.qubits, args.depth
    np_dtype = np.complex64 if args.dtype == 'c64' else np.complex128
    amp_bytes = 8 if args.dtype == 'c64' else 16
    t0 = time.perf_counter()
    qpe = None
    if args.workload == 'qpe':
        from qcircuits_b200.workloads import phase_estimation_circuit
        circ, eigvec, phase = phase_estimation_circuit(n - 2, seed=SEED)
        qpe = {'true_phase': phase}
    else:
        circ = layered_circuit(n, depth, seed=SEED)
    fuse = not args.no_fuse
    passes, ngates = circ.stats(np_dtype, fuse=fuse, max_ops=args.max_ops)
    plan_s = time.perf_counter() - t0

    if qpe is None:
        state0 = qc.zeros(n, dtype=np_dtype)
    else:
        state0 = qc.zeros(n - 2, dtype=np_dtype) * qc.State(eigvec.reshape(2, 2), dtype=np_dtype)
    for _ in range(args.warmup):
        out = circ.run(state0, fuse=fuse, max_ops=args.max_ops)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(0) as clocks:
        torch.cuda.synchronize()
        ev[0].record()
        for s in range(args.steps):
            out = circ.run(state0, fuse=fuse, max_ops=args.max_ops)
            ev[s + 1].record()
        torch.cuda.synchronize()
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    ms_per_step = total_ms / args.steps
    value = ngates / (ms_per_step * 1e-3)
    pass_bytes = 2.0 * (2 ** n) * amp_bytes
    achieved = pass_bytes * passes / (ms_per_step * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    # the per-gate path (Operator.__call__: one pass per gate, north_star "per gate pass"): the first two
    # layers of the same circuit, unfused, timed the same way
    if qpe is not None:
        k = min(12, n - 2)
        ps = out._measurement_probabilities(list(range(k)))
        qpe.update({'estimated_phase_%d_bits' % k: int(np.argmax(ps)) / 2.0 ** k, 'peak_probability': float(ps.max())})
    from qcircuits_b200.workloads import layered_gate_list, operator_for
    from qcircuits_b200.circuit import Circuit
    two_layers = Circuit(n)
    for name, theta, qs in layered_gate_list(n, 2, seed=SEED):
        two_layers.append(operator_for(name, theta), list(qs))
    for _ in range(2):
        two_layers.run(state0, fuse=False)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(3):
        two_layers.run(state0, fuse=False)
    g1.record()
    torch.cuda.synchronize()
    gate_ms = g0.elapsed_time(g1) / (3 * len(two_layers))
    per_gate = {'bound': 'hbm', 'achieved': pass_bytes / (gate_ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                'frac': pass_bytes / (gate_ms * 1e-3) / 1e9 / peak, 'ms_per_launch': gate_ms,
                'gates_per_s': 1e3 / gate_ms, 'launches_timed': 3 * len(two_layers),
                'frac_of_8TBs_nominal': pass_bytes / (gate_ms * 1e-3) / 1e9 / 8000.0,
                'what': 'one tile-kernel pass per gate (first two layers of the circuit, unfused)'}
    # sanity: unit norm after 1780 gates
    probs_sum = float(out._dv.norm2()[0].item()) / float(out._dv.norm[0].item())

    e2e = None
    if not args.skip_e2e:
        real = torch.float32 if args.dtype == 'c64' else torch.float64
        host_in = torch.zeros(2 << n, dtype=real).pin_memory()
        host_in[0] = 1.0
        host_out = torch.empty(2 << n, dtype=real).pin_memory()
        for _ in range(max(1, min(args.warmup, 3))):
            circ.run_host(host_in, host_out, fuse=fuse, max_ops=args.max_ops)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(args.steps):
            circ.run_host(host_in, host_out, fuse=fuse, max_ops=args.max_ops)
        e1.record()
        torch.cuda.synchronize()
        e2e_ms = e0.elapsed_time(e1) / args.steps
        nrm = float((host_out.double() ** 2).sum())
        e2e = {'value': ngates / (e2e_ms * 1e-3), 'unit': 'gates/s', 'h2d_bytes_per_step': int(2 ** n * amp_bytes),
               'd2h_bytes_per_step': int(2 ** n * amp_bytes), 'ms_per_step': e2e_ms, 'result_norm2': nrm}
        del host_in, host_out

    cpu = None
    if not args.skip_cpu:
        from oracle import qc_oracle as orc      # the cpu_baseline leg: the only use of oracle/ here
        n_sample = min(args.cpu_qubits, n)
        scaled, raw, done, dt = cpu_port_gates_per_s(orc, n_sample, n, args.cpu_budget, depth)
        cpu = {'value': scaled, 'unit': 'gates/s', 'cores': 1, 'kind': 'port', 'host_cores': os.cpu_count(),
               'sample': '%d gates of the same circuit at %d qubits complex128 in %.1f s (%.3f gates/s), scaled by '
                         '2^(%d-%d)' % (done, n_sample, dt, raw, n_sample, n)}

    line = {
        'metric': 'gates/s, random layered circuit', 'value': value, 'unit': 'gates/s', 'n_gpus': 1,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': ('layered H/RX/RZ+CNOT/CZ circuit, %d qubits, depth %d, seed %d, |0..0> input' % (n, depth, SEED)
                                if qpe is None else
                                'phase estimation of a seeded 2-qubit unitary, %d-qubit register + inverse QFT, %d qubits, seed %d'
                                % (n - 2, n, SEED)),
                   'gates': ngates, 'passes_per_step': passes, 'fused': fuse, 'max_ops_per_pass': args.max_ops,
                   'l2': 'state (%.1f GiB) is far larger than L2; no flush needed' % (2 ** n * amp_bytes / 2 ** 30),
                   'plan_s': plan_s},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': recorded_traffic('tile_%s_%dq' % (args.dtype, n)), 'peak_source': peak_src,
                     'kernel': 'qcs::tile_kernel (fused gate pass)', 'bytes_per_launch': pass_bytes,
                     'ms_per_launch': ms_per_step / passes, 'frac_of_8TBs_nominal': achieved / 8000.0},
        'roofline_per_gate_pass': per_gate,
        'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': passes * args.steps,
        'clocks': clocks.summary(), 'step_ms': step_ms, 'norm_check': probs_sum,
    }
    if qpe is not None:
        line['qpe_check'] = qpe
        line['metric'] = 'gates/s, phase estimation + inverse QFT'
    line['roofline']['gate_equivalent_GBs'] = pass_bytes * ngates / (ms_per_step * 1e-3) / 1e9
    print(json.dumps(line))


if __name__ == '__main__':
    main()
