"""Shared helpers for the parity tests (CPU-side generators, oracle adapters)."""
import numpy as np

from oracle import qc_oracle as orc


def rand_state_vec(rng, n):
    v = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    return v / np.linalg.norm(v)


def rand_unitary(rng, k):
    A = rng.normal(size=(2 ** k, 2 ** k)) + 1j * rng.normal(size=(2 ** k, 2 ** k))
    Q, R = np.linalg.qr(A)
    return Q * (np.diag(R) / np.abs(np.diag(R)))


def oracle_apply(M, vec, qubit_indices, renorm=True):
    n = int(round(np.log2(vec.size)))
    return orc.apply_matrix_flat(M, vec, [n - 1 - q for q in qubit_indices], renorm=renorm)


def rel_err(got, want):
    return float(np.max(np.abs(np.asarray(got) - np.asarray(want))) / max(np.max(np.abs(want)), 1e-300))


TOL = {np.dtype(np.complex128): 1e-12, np.dtype(np.complex64): 1e-5}


# achieved error / tolerance of every parity assertion made through check_close, reported at the end of the run
# (tests/conftest.py) so that the log shows how far each test is from the north-star bar, not only that it passed
ACHIEVED = {}


def check_close(got, want, dtype, label, factor=1.0):
    """assert rel_err(got, want) < factor * TOL[dtype] and record err / TOL.  `factor` != 1 must be justified at the
    call site with the measured number."""
    tol = TOL[np.dtype(dtype)]
    err = rel_err(got, want)
    key = '%s[%s]' % (label, np.dtype(dtype).name)
    ACHIEVED[key] = max(ACHIEVED.get(key, 0.0), err / tol)
    assert err < factor * tol, '%s: error %.3e = %.2f x the %.0e tolerance' % (key, err, err / tol, tol)
    return err
