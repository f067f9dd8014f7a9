"""The oracle restatement against the UNMODIFIED reference vendored as oracle/_ref (oracle/make_ref.py) -- run here
and on the GPU box, where /root/reference does not exist but oracle/_ref has travelled with the snapshot.  Together
with tests/test_oracle_golden.py (committed golden vectors of the same reference) this pins the oracle."""
import hashlib
import os

import numpy as np
import pytest

from oracle import make_ref
from oracle import qc_oracle as orc
from tests.util import rand_state_vec, rand_unitary

ref = make_ref.load_reference()
pytestmark = pytest.mark.skipif(ref is None, reason='oracle/_ref has not been made (python oracle/make_ref.py)')


def test_the_vendored_reference_is_byte_identical_to_its_manifest():
    root = os.path.join(os.path.dirname(make_ref.__file__), '_ref')
    lines = open(os.path.join(root, 'MANIFEST.sha256')).read().split('\n')
    assert len([l for l in lines if l]) >= 15
    for line in lines:
        if not line:
            continue
        digest, rel = line.split('  ')
        assert hashlib.sha256(open(os.path.join(root, rel), 'rb').read()).hexdigest() == digest, rel
    if os.path.isdir('/root/reference/qcircuits'):           # in the build container: identical to the source tree
        for line in lines:
            if line:
                digest, rel = line.split('  ')
                assert hashlib.sha256(open(os.path.join('/root/reference', rel), 'rb').read()).hexdigest() == digest, rel


def test_gate_application_matches_the_reference():
    rng = np.random.default_rng(100)
    for n in (1, 2, 3, 5, 8, 11):
        vec = rand_state_vec(rng, n)
        for k in range(1, min(n, 4) + 1):
            qi = [int(q) for q in rng.permutation(n)[:k]]
            U = rand_unitary(rng, k)
            want = ref.Operator.from_matrix(U)(ref.State(vec.reshape([2] * n)), qubit_indices=qi).to_column_vector()
            got = orc.apply_matrix_flat(U, vec, [n - 1 - q for q in qi], renorm=True)
            assert np.max(np.abs(got - want)) < 1e-13, (n, qi)
        # non-unitary operators are renormalised the same way (operators.py:275-276)
        A = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
        if n >= 2:
            qi = [int(q) for q in rng.permutation(n)[:2]]
            want = ref.Operator.from_matrix(A)(ref.State(vec.reshape([2] * n)), qubit_indices=qi).to_column_vector()
            got = orc.apply_matrix_flat(A, vec, [n - 1 - q for q in qi], renorm=True)
            assert np.max(np.abs(got - want)) < 1e-12


def test_layered_circuit_and_measurement_match_the_reference():
    n, depth = 9, 6
    gates = orc.layered_circuit(n, depth, seed=1234)
    mk = {'H': lambda t: ref.Hadamard(), 'RX': ref.RotationX, 'RZ': ref.RotationZ, 'CNOT': lambda t: ref.CNOT(),
          'CZ': lambda t: ref.ControlledU(ref.PauliZ())}
    state = ref.zeros(n)
    for name, theta, qs in gates:
        state = mk[name](theta)(state, qubit_indices=list(qs))
    want = state.to_column_vector()
    got = orc.run_circuit(orc.zeros_state(n), gates).reshape(-1)
    assert np.max(np.abs(got - want)) < 1e-13
    assert np.max(np.abs(orc.probabilities(got.reshape([2] * n)) - state.probabilities)) < 1e-13
    # measurement: same uniform -> same outcome and post-measurement state (np.random.choice draws exactly one)
    for seed, qs, remove in ((1, [3, 0, 7], False), (2, [8], True), (3, [1, 2, 3, 4], True), (4, None, False)):
        s2 = ref.State(want.reshape([2] * n))
        np.random.seed(seed)
        bits_ref = s2.measure(qubit_indices=qs, remove=remove)
        np.random.seed(seed)
        u = np.random.random_sample()
        bits, post = orc.measure(want.reshape([2] * n), list(range(n)) if qs is None else qs, u, remove=remove)
        assert tuple(np.atleast_1d(bits_ref)) == tuple(bits)
        assert np.max(np.abs(post.reshape(-1) - s2.to_column_vector())) < 1e-13


def test_density_operator_paths_match_the_reference():
    rng = np.random.default_rng(101)
    n = 4
    with make_ref.reference_names():
        states = [ref.State(rand_state_vec(rng, n).reshape([2] * n)) for _ in range(3)]
        ps = [0.5, 0.3, 0.2]
        rho = ref.DensityOperator.from_ensemble(states, ps)
        U = rand_unitary(rng, 2)
        out = ref.Operator.from_matrix(U)(rho, qubit_indices=[2, 0])
        want_t, want_p = np.array(out._t), np.array(out.probabilities)
        red = np.array(out.reduced_density_operator([1, 3])._t)
    t = orc.from_ensemble([np.array(s._t) for s in states], ps)
    t = orc.apply_operator_density(orc.op_from_matrix(U), t, [2, 0])
    assert np.max(np.abs(t - want_t)) < 1e-13
    assert np.max(np.abs(orc.density_probabilities(t) - want_p)) < 1e-13
    assert np.max(np.abs(orc.reduced_tensor(t, [1, 3]) - red)) < 1e-13
    # measurement of the mixed state: same uniform -> same outcome and post-measurement operator
    for seed, qs, remove in ((5, [1, 3], False), (6, [0], True)):
        with make_ref.reference_names():
            r2 = ref.DensityOperator(want_t)
            np.random.seed(seed)
            bits_ref = r2.measure(qubit_indices=qs, remove=remove)
            post_ref = np.array(r2._t)
        np.random.seed(seed)
        bits, post = orc.density_measure(want_t, qs, np.random.random_sample(), remove=remove)
        assert tuple(np.atleast_1d(bits_ref)) == tuple(bits)
        assert np.max(np.abs(post - post_ref)) < 1e-13
