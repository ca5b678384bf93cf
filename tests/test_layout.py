"""Repo-level invariants the task contract names."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(d):
    for base, _, files in os.walk(os.path.join(ROOT, d)):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(base, f)


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle/|liboracle", re.M)
    for p in _py_files("ffwm_b200"):
        assert not pat.search(open(p).read()), "%s references the oracle" % p


def test_oracle_files_say_they_are_test_infrastructure():
    for f in ("warp_ops.c", "warp_ops.inc", "warp.py", "build_ref.py", "__init__.py"):
        assert "test infrastructure" in open(os.path.join(ROOT, "oracle", f)).read()


def test_no_reference_reads_at_runtime():
    # /root/reference does not exist on the GPU box: product, bench and smoke must not read it
    for p in list(_py_files("ffwm_b200")) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]:
        if not os.path.exists(p):
            continue
        text = open(p).read()
        assert "/root/reference" not in text.replace("/root/reference is", "").replace("(/root/reference", ""), p
