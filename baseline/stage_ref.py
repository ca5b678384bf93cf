"""Stage the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box).

The reference (csyxwei/FFWM) is a script repository without setup.py / pyproject, so
`pip install --target baseline/_ref /root/reference` has nothing to install; the equivalent is a
byte-for-byte copy of the two Python packages the orchestrators import (`models/`, `lightcnn/`).
Nothing is edited: `baseline/_ref/MANIFEST.json` records the sha256 of every staged file next to
the sha256 of its source, and bench.py / the tests apply harness-side shims only (SURVEY 8c:
`numpy.int`, an offline VGG19 cache file, random state_dicts at the `opt.*` checkpoint paths).

Used by: `bench.py --impl reference` (the reference's own CPU path, FFWMModel(gpu_ids=[])),
tests/test_orchestrators_gpu.py (the unmodified orchestrators stepping on the sm_100a kernels),
tests/golden/make_golden_*.py.  Never imported by ffwm_b200/.

    python baseline/stage_ref.py            # no-op where /root/reference is absent
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FFWM_REFERENCE", os.path.join(os.sep, "root", "reference"))
DST = os.path.join(HERE, "_ref")
PACKAGES = ("models", "lightcnn")


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    if not os.path.isdir(REF):
        print("stage_ref: %s not present (GPU box uses the staged copy)" % REF)
        return 0
    manifest = {}
    for pkg in PACKAGES:
        for dirpath, _, files in os.walk(os.path.join(REF, pkg)):
            for f in files:
                if not f.endswith(".py"):
                    continue
                src = os.path.join(dirpath, f)
                rel = os.path.relpath(src, REF)
                dst = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                manifest[rel] = {"sha256": sha(dst), "source_sha256": sha(src)}
                assert manifest[rel]["sha256"] == manifest[rel]["source_sha256"]
    json.dump({"source": REF, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    print("stage_ref: %d files -> %s" % (len(manifest), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
