"""Recipe for oracle/_ref: the UNMODIFIED reference, copied from where it lies under /root/reference.

TEST INFRASTRUCTURE, not product: only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs
of bench.py may import or execute anything under oracle/.

The reference (grey-area/qcircuits v0.6.0) is four pure-Python files; there is nothing to compile, so the
"build" is a byte-for-byte copy of
    /root/reference/qcircuits/*.py            -> oracle/_ref/qcircuits/         (the implementation the oracle restates)
    /root/reference/tests/{tests,slow_measurement_tests}.py -> oracle/_ref/tests/   (its own property suite)
    /root/reference/examples/*.py             -> oracle/_ref/examples/          (Grover, phase estimation, ...)
plus a MANIFEST of sha256 digests.  oracle/_ref/ is git-ignored (no reference source enters the history) but is
NOT gpurun-ignored: it travels to the GPU box with the snapshot, where /root/reference does not exist.  There it
serves (1) as the checker the oracle restatement is validated against, (2) as the CPU baseline that bench.py times
(`cpu_baseline.kind = "reference"`), (3) as the unmodified test-suite and examples that are run on the CUDA engine.

Run:  python oracle/make_ref.py            (also called by __graft_entry__.build() when /root/reference is present)
"""
import hashlib
import os
import shutil
import sys

SRC = os.environ.get('QCS_REFERENCE_DIR', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')

COPIES = [('qcircuits', ['__init__.py', 'tensors.py', 'state.py', 'operators.py', 'density_operator.py']),
          ('tests', ['tests.py', 'slow_measurement_tests.py']),
          ('examples', None)]          # None: every .py file of the directory


def make():
    if not os.path.isdir(os.path.join(SRC, 'qcircuits')):
        return False
    manifest = []
    for sub, names in COPIES:
        src_dir, dst_dir = os.path.join(SRC, sub), os.path.join(DST, sub)
        os.makedirs(dst_dir, exist_ok=True)
        if names is None:
            names = sorted(f for f in os.listdir(src_dir) if f.endswith('.py'))
        for name in names:
            shutil.copyfile(os.path.join(src_dir, name), os.path.join(dst_dir, name))
            with open(os.path.join(dst_dir, name), 'rb') as f:
                manifest.append('%s  %s/%s' % (hashlib.sha256(f.read()).hexdigest(), sub, name))
    with open(os.path.join(DST, 'MANIFEST.sha256'), 'w') as f:
        f.write('\n'.join(manifest) + '\n')
    return True


def load_reference():
    """Import the vendored reference as a module named ``qcircuits_ref`` (never as ``qcircuits``: that name is what
    install_as_qcircuits() binds to the CUDA engine).  Returns None when oracle/_ref has not been made."""
    import importlib.util
    pkg_dir = os.path.join(DST, 'qcircuits')
    if not os.path.isfile(os.path.join(pkg_dir, '__init__.py')):
        return None
    if 'qcircuits_ref' in sys.modules:
        return sys.modules['qcircuits_ref']
    import warnings
    spec = importlib.util.spec_from_file_location('qcircuits_ref', os.path.join(pkg_dir, '__init__.py'),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules['qcircuits_ref'] = mod
    # the reference imports its siblings absolutely ("from qcircuits.tensors import Tensor"): alias them for the
    # duration of the import, then restore whatever was registered under those names
    saved = {k: v for k, v in sys.modules.items() if k == 'qcircuits' or k.startswith('qcircuits.')}
    for k in saved:
        del sys.modules[k]
    sys.modules['qcircuits'] = mod
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')           # invalid escape sequences in the reference's docstrings
            spec.loader.exec_module(mod)
        for k in [k for k in sys.modules if k.startswith('qcircuits.')]:
            sys.modules['qcircuits_ref.' + k.split('.', 1)[1]] = sys.modules.pop(k)
    finally:
        sys.modules.pop('qcircuits', None)
        sys.modules.update(saved)
    return mod


class reference_names:
    """``with reference_names():`` -- while inside, ``qcircuits`` and ``qcircuits.*`` name the vendored reference.
    Needed around reference calls that import lazily at call time (State.density_operator, DensityOperator.measure,
    purify: reference state.py:281,302, density_operator.py:163,228)."""

    def __enter__(self):
        ref = load_reference()
        self.saved = {k: v for k, v in sys.modules.items() if k == 'qcircuits' or k.startswith('qcircuits.')}
        for k in self.saved:
            del sys.modules[k]
        sys.modules['qcircuits'] = ref
        for k, v in list(sys.modules.items()):
            if k.startswith('qcircuits_ref.'):
                sys.modules['qcircuits.' + k.split('.', 1)[1]] = v
        return ref

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k == 'qcircuits' or k.startswith('qcircuits.')]:
            del sys.modules[k]
        sys.modules.update(self.saved)
        return False


if __name__ == '__main__':
    ok = make()
    print('oracle/_ref made from %s' % SRC if ok else 'no reference at %s: oracle/_ref left as it is' % SRC)
