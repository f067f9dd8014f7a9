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
  cpu_baseline   the UNMODIFIED reference (oracle/_ref, kind "reference"; copied by oracle/make_ref.py, it travels to
             the GPU box) on the host cores: measured at 20/22/24/26 qubits, 2^n fit, extrapolated to the workload

`--impl reference` times only that reference CPU path, same metric and config.
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


def _cpu_engine():
    """(kind, run_gates): the reference's own CPU implementation of the path.  kind "reference" = the UNMODIFIED
    reference vendored as oracle/_ref by oracle/make_ref.py (it travels to the GPU box with the snapshot); kind "port" =
    the oracle's restatement of the same route (tensordot + copy + 4-pass renormalise), only when oracle/_ref is absent.
    run_gates(n, gates, budget_s) applies the gate list to |0..0> until the budget is spent (the first gate -- first
    touch of the arrays -- is excluded) and returns (gates timed, seconds)."""
    from oracle import make_ref
    ref = make_ref.load_reference()
    if ref is not None:
        mk = {'H': lambda t: ref.Hadamard(), 'RX': ref.RotationX, 'RZ': ref.RotationZ, 'CNOT': lambda t: ref.CNOT(),
              'CZ': lambda t: ref.ControlledU(ref.PauliZ())}

        def run_gates(n, gates, budget_s):
            state = ref.zeros(n)
            state = mk[gates[0][0]](gates[0][1])(state, qubit_indices=list(gates[0][2]))
            done, t0 = 0, time.perf_counter()
            for name, theta, qs in gates[1:]:
                state = mk[name](theta)(state, qubit_indices=list(qs))
                done += 1
                if time.perf_counter() - t0 > budget_s:
                    break
            return done, time.perf_counter() - t0
        return 'reference', run_gates
    from oracle import qc_oracle as orc

    def run_gates(n, gates, budget_s):
        psi = orc.zeros_state(n)
        psi = orc.apply_operator(orc.gate_tensor(*gates[0][:2]), psi, list(gates[0][2]))
        done, t0 = 0, time.perf_counter()
        for name, theta, qs in gates[1:]:
            psi = orc.apply_operator(orc.gate_tensor(name, theta), psi, list(qs))
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        return done, time.perf_counter() - t0
    return 'port', run_gates


def _blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([int(p.get('num_threads', 1)) for p in threadpool_info()] or [1])
    except Exception:
        return None


def cpu_reference_rate(n_target, depth, budget_s, sizes=None):
    """gates/s of the reference CPU path on the layered circuit: MEASURED at several sizes that fit the time budget
    (whole layers of the n-qubit instance of the same generator), a 2^(slope * n) fit through them, and the
    EXTRAPOLATION to n_target when that size cannot be run (30 qubits complex128 needs 4 x 16 GiB and ~40 s per gate).
    Returns a cpu_baseline record."""
    from qcircuits_b200.workloads import layered_gate_list
    kind, run_gates = _cpu_engine()
    sizes = sizes or [s for s in (20, 22, 24, 26) if s <= n_target]
    share = budget_s / max(1, len(sizes))
    points = []
    for n in sizes:
        gates = layered_gate_list(n, depth, seed=SEED)
        done, dt = run_gates(n, gates, share)
        points.append({'qubits': n, 'gates_timed': done, 'seconds': dt, 'gates_per_s': done / dt})
    xs = np.array([p['qubits'] for p in points], dtype=float)
    ys = np.log2([p['gates_per_s'] for p in points])
    slope = float(np.polyfit(xs, ys, 1)[0]) if len(points) > 1 else -1.0
    last = points[-1]
    measured_at_target = last['qubits'] == n_target
    value = last['gates_per_s'] * 2.0 ** (slope * (n_target - last['qubits']))
    return {'value': value, 'unit': 'gates/s', 'cores': 1, 'kind': kind, 'host_cores': os.cpu_count(),
            'blas_threads': _blas_threads(), 'extrapolated': not measured_at_target,
            'fit': {'log2_gates_per_s_per_qubit': slope, 'points': points},
            'sample': ('%s path, complex128 (the reference has no complex64), whole gates of the %d-layer circuit at %s '
                       'qubits for %.0f s each, first-touch gate excluded; %s to %d qubits with the fitted slope %.2f '
                       'bits/qubit; the path is effectively single-threaded (cores = 1)'
                       % ('unmodified reference (oracle/_ref)' if kind == 'reference' else 'oracle port of the reference route',
                          depth, '/'.join(str(p['qubits']) for p in points), share,
                          'measured at' if measured_at_target else 'extrapolated', n_target, slope))}


def reference_arm(args):
    """`bench.py --impl reference`: the reference's own CPU implementation of the path on the box's host cores, same
    metric / config as the CUDA arm; every step is a bounded sample (see cpu_reference_rate)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = args.qubits
    per_step = []
    rec = None
    for _ in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rec = cpu_reference_rate(n, args.depth, args.cpu_budget)
        per_step.append((rec['value'], time.perf_counter() - t0))
    timed = per_step[args.warmup:]
    value = float(np.mean([v for v, _ in timed]))
    rec = dict(rec, value=value)
    line = {'impl': 'reference', 'metric': 'gates/s, random layered circuit', 'value': value, 'unit': 'gates/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * float(np.mean([d for _, d in timed])), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'c128', 'data': 'synthetic',
            'config': {'workload': 'layered H/RX/RZ+CNOT/CZ circuit, %d qubits, depth %d, seed %d, |0..0> input' % (n, args.depth, SEED)},
            'cpu_baseline': rec,
            'e2e': {'value': value, 'unit': 'gates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def sharded_parity_check():
    """Before the multi-GPU timing: ShardedState (circuit with exchanges, Operator.__call__, marginals, measure, tensor
    product) against the ORACLE at 14 and 19 qubits on this very process group (tools/dist_check.py; the oracle is the
    checker here, never the thing measured).  Returns the record that goes into the line as `parity`."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('dist_check', os.path.join(ROOT, 'tools', 'dist_check.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    worst = mod.check(max_n=19, verbose=False)
    return {'checked_against': 'oracle/qc_oracle.py', 'qubits': [14, 19], 'dtypes': ['c64', 'c128'],
            'worst_error_over_tolerance': worst, 'tolerance': {'c64': 1e-5, 'c128': 1e-12}, 'ok': bool(worst < 1.0)}


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
    ap.add_argument('--cpu-budget', type=float, default=24.0, help='seconds of CPU work per reference sample')
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--skip-e2e', action='store_true')
    ap.add_argument('--skip-strong', action='store_true', help='skip the 33-qubit strong-scaling leg')
    ap.add_argument('--skip-configs', action='store_true', help='skip the short runs of BASELINE configs 2, 3 and 5')
    ap.add_argument('--strong-qubits', type=int, default=33)
    ap.add_argument('--skip-parity', action='store_true', help='multi-GPU: skip the sharded-vs-oracle check before timing')
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
        return dist_bench.main(args, parity_check=None if args.skip_parity else sharded_parity_check)

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
        ps = out.marginal_probabilities(list(range(k)))
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
    if not args.skip_cpu:        # the cpu_baseline leg: the only use of oracle/ here
        cpu = cpu_reference_rate(n, depth, args.cpu_budget)

    # ---- the other BASELINE configs and the strong-scaling leg, each a short device-resident run ----------------
    del out, state0
    torch.cuda.empty_cache()

    def timed_run(circuit, make_state, reps, dtype):
        s0 = make_state()
        res = circuit.run(s0)                                 # warm-up (plans, first touch)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            res = None                                        # (never two results alive: 33 qubits = 64 GiB each)
            res = circuit.run(s0)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        npass, ng = circuit.stats(dtype)
        return res, {'ms_per_step': ms, 'gates': ng, 'passes': npass, 'gates_per_s': ng / (ms * 1e-3)}

    strong = None
    if not args.skip_strong and qpe is None and args.dtype == 'c64' and torch.cuda.mem_get_info()[0] > 140 * 2 ** 30:
        ns = args.strong_qubits
        cs = layered_circuit(ns, depth, seed=SEED)
        res, strong = timed_run(cs, lambda: qc.zeros(ns, dtype=np.complex64), 2, np.complex64)
        strong.update({'qubits': ns, 'amp_updates_per_s_per_gpu': strong['gates'] * 2.0 ** ns / (strong['ms_per_step'] * 1e-3),
                       'what': 'the same 33-qubit circuit the multi-GPU lines run sharded (strong scaling, N = 1 leg)'})
        del res, cs
        torch.cuda.empty_cache()
    configs = None
    if not args.skip_configs and qpe is None and args.dtype == 'c64' and n == 30:
        configs = {}
        c2 = layered_circuit(28, 40, seed=SEED)
        res, rec = timed_run(c2, lambda: qc.zeros(28, dtype=np.complex128), 3, np.complex128)
        rec.update({'workload': 'config 2: layered circuit, 28 qubits complex128, depth 40',
                    'hbm_GBs': 2.0 * 2 ** 28 * 16 * rec['passes'] / (rec['ms_per_step'] * 1e-3) / 1e9})
        configs['config2_28q_c128'] = rec
        del res, c2
        torch.cuda.empty_cache()
        if torch.cuda.mem_get_info()[0] > 70 * 2 ** 30:
            from qcircuits_b200.workloads import phase_estimation_circuit
            c3, eigvec, phase = phase_estimation_circuit(30, seed=SEED)
            res, rec = timed_run(c3, lambda: qc.zeros(30, dtype=np.complex64) * qc.State(eigvec.reshape(2, 2), dtype=np.complex64),
                                 1, np.complex64)
            ps = res.marginal_probabilities(list(range(12)))
            rec.update({'workload': 'config 3: phase estimation of a seeded 2-qubit unitary, t = 30, 32 qubits complex64',
                        'true_phase': phase, 'estimated_phase_12_bits': int(np.argmax(ps)) / 4096.0,
                        'peak_probability': float(ps.max())})
            configs['config3_qpe_32q_c64'] = rec
            del res, c3
            torch.cuda.empty_cache()
        from qcircuits_b200.workloads import layered_gate_list, operator_for
        rho = qc.zeros(14).density_operator()
        glist = [(operator_for(nm, th), list(qs)) for nm, th, qs in layered_gate_list(14, 10, seed=SEED)]
        for op, qs in glist[:20]:
            rho = op(rho, qubit_indices=qs)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for op, qs in glist[20:]:
            rho = op(rho, qubit_indices=qs)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        configs['config5_rho_14q_c128'] = {
            'workload': 'config 5: U rho U^dagger gate by gate (Operator.__call__), 14 qubits = rank-28 tensor complex128, depth 10',
            'gates': len(glist) - 20, 'ms_total': ms, 'gates_per_s': (len(glist) - 20) / (ms * 1e-3),
            'hbm_GBs': 2.0 * 2 ** 28 * 16 * (len(glist) - 20) / (ms * 1e-3) / 1e9, 'trace': float(rho.probabilities.sum())}
        del rho
        torch.cuda.empty_cache()

    # how the passes run, and what the tensor-core blocks amount to: a block applies 18 tcgen05.mma kind::f16 per 128-row
    # tile (6 x M128 N64 K16 + 12 x M128 N32 K16: hi.hi + hi.lo + lo.hi of the two-term fp16 split) = 3.15 MFLOP per
    # tile of 2^12 amplitudes
    plan = circ.plan_summary(np_dtype, fuse=fuse, max_ops=args.max_ops)
    tensor = None
    if plan['blocks']:
        flop = plan['blocks'] * (2.0 ** n / 4096.0) * (6 * 128 * 64 * 16 * 2 + 12 * 128 * 32 * 16 * 2)
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                f16_peak = float(json.load(f)['bf16_tflops_sustained'])
            f16_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (fp16 runs at the bf16 rate)'
        except Exception:
            f16_peak, f16_src = 1400.0, 'B200_PROFILING.md fallback 1.4 PFLOP/s sustained bf16'
        tensor = {'f16_tflops_achieved_over_the_step': flop / (ms_per_step * 1e-3) / 1e12, 'f16_peak_tflops': f16_peak,
                  'frac': flop / (ms_per_step * 1e-3) / 1e12 / f16_peak, 'peak_source': f16_src,
                  'note': 'executed fp16 flops (3 terms of the split) of the dense blocks / step time; the pass is bound by the '
                          'serial chain of a sweep and the shared-memory pipe, not the tensor pipe (DESIGN.md 3.1b)'}

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
                     'traffic': recorded_traffic('tile_%s_%dq' % (args.dtype, n)),
                     'traffic_source': 'profiles/traffic.json: dram__bytes read + written per launch of the committed ncu --set full '
                                       'capture of the same kernel (a recorded constant, not measured in this run)',
                     'peak_source': peak_src,
                     'kernel': ('qcs::tile_kernel_tc (fused gate pass: dense 2^5 blocks on tcgen05 + register ops)'
                                if plan['tensor_core_passes'] else 'qcs::tile_kernel_pipe (fused gate pass, register sweeps)'),
                     'bytes_per_launch': pass_bytes, 'ms_per_launch': ms_per_step / passes,
                     'frac_of_8TBs_nominal': achieved / 8000.0, 'plan': plan, 'tensor_pipe': tensor},
        'roofline_per_gate_pass': per_gate,
        'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': passes * args.steps,
        'amp_updates_per_s_per_gpu': ngates * 2.0 ** n / (ms_per_step * 1e-3),
        'strong_scaling': strong, 'other_configs': configs,
        'clocks': clocks.summary(), 'step_ms': step_ms, 'norm_check': probs_sum,
    }
    if qpe is not None:
        line['qpe_check'] = qpe
        line['metric'] = 'gates/s, phase estimation + inverse QFT'
    line['roofline']['gate_equivalent_GBs'] = pass_bytes * ngates / (ms_per_step * 1e-3) / 1e9
    print(json.dumps(line))


if __name__ == '__main__':
    main()
