"""bench.py and tools/*.py only ever run on the GPU box; a name that is bound in a different function
would surface there as a NameError after minutes of set-up.  tools/check_names.py (a scope-aware
undefined-name pass; no pyflakes in the image) runs over them here."""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_no_undefined_names():
    import check_names
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    files += sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))) + sorted(glob.glob(os.path.join(ROOT, "monocle3_b200", "*.py")))
    bad = [(f, line, name) for f in files for line, name in check_names.check(f)]
    assert not bad, bad
