"""
TEST INFRASTRUCTURE ONLY -- stages the reference's own test-suite for a run
against the drop-in package on the GPU box.

BASELINE.md / SURVEY 4(i): the 74 tests of ``test_fcdiff/`` (test_fit.py 50,
test_model.py 21, test_util.py 3) are a parity gate for ``import fcdiff``.
They need a GPU (every ``fcdiff.fit`` / ``fcdiff.model`` call runs CUDA) and
``/root/reference`` does not exist on the GPU box, so -- exactly like a compiled
``oracle/_ref`` artefact -- the test files are COPIED, byte for byte, into
``oracle/_ref/ref_tests/test_fcdiff/`` (git-ignored: no reference source enters
the history; not gpurun-ignored: the copy travels with the snapshot).  The only
thing added is a ``conftest.py`` of ours beside them with the one shim the tests
need on a current SciPy: ``scipy.misc.logsumexp`` (removed in SciPy 1.3; P4 of
SURVEY 0.2; used at test_fit.py:465, 507) -> ``scipy.special.logsumexp``.

    python oracle/stage_ref_tests.py        # also run by __graft_entry__.build()

``tests/test_reference_suite.py`` (``-m gpu``) then runs pytest on the staged
directory with the repo root on ``sys.path``, so that ``import fcdiff`` resolves
to the alias package of this repo.
"""
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_TESTS = "/root/reference/test_fcdiff"
STAGED = os.path.join(ROOT, "oracle", "_ref", "ref_tests")

CONFTEST = '''"""Shim for the reference's unmodified tests (written by oracle/stage_ref_tests.py)."""
import os
import sys
import warnings

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    import scipy.misc
    import scipy.special
if not hasattr(scipy.misc, "logsumexp"):              # P4 of SURVEY 0.2
    scipy.misc.logsumexp = scipy.special.logsumexp
'''


def stage(force=False):
    """Copies the reference's test files (unmodified) when /root/reference is present.
    Returns the staged directory, or None when there is nothing to stage from and
    nothing staged earlier."""
    dst = os.path.join(STAGED, "test_fcdiff")
    if os.path.isdir(REFERENCE_TESTS):
        os.makedirs(dst, exist_ok=True)
        for name in sorted(os.listdir(REFERENCE_TESTS)):
            if name.endswith(".py"):
                src = os.path.join(REFERENCE_TESTS, name)
                out = os.path.join(dst, name)
                if force or not os.path.isfile(out) or os.path.getmtime(out) < os.path.getmtime(src):
                    shutil.copyfile(src, out)
        with open(os.path.join(STAGED, "conftest.py"), "w") as f:
            f.write(CONFTEST)
        with open(os.path.join(STAGED, "pytest.ini"), "w") as f:      # its own rootdir: no markers, no testpaths
            f.write("[pytest]\nfilterwarnings =\n    ignore::DeprecationWarning\n    ignore::SyntaxWarning\n")
    return STAGED if os.path.isdir(dst) else None


if __name__ == "__main__":
    print(stage(force=True))
