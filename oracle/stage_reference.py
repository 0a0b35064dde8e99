"""Stage the reference's own Python sources for the hot path into oracle/_ref/  --  TEST INFRASTRUCTURE.

The reference (nicholasburden/pymarl) is a plain source tree (no setup.py, nothing to compile): its learner is
`src/learners/q_learner.py` on top of torch.  The GPU box has torch but no /root/reference, so `build()` (run in the
build container) copies the reference's `.py` / `.yaml` files - unmodified, byte for byte - from where they lie into
the git-ignored `oracle/_ref/src/` (listed in .gitignore, NOT in .gpurunignore: it travels to the GPU box like a built
.so, and never enters the history).  `bench.py --impl reference` and the `cpu_baseline` leg then time the REAL
`QLearner.train(use_cuda=False)` / `BasicMAC.select_actions` on the box's host cores (`kind: "reference"`); tests use it
to drive the reference's `run.run_sequential` with this package's classes injected.

    python oracle/stage_reference.py [--src /root/reference/src]

Nothing under pymarl_b200/ imports oracle/ (the product has no CPU path).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "src")
DEFAULT_SRC = os.environ.get("PYMARL_REF_SRC", "/root/reference/src")
KEEP_EXT = (".py", ".yaml")


def stage(src=DEFAULT_SRC, dest=DEST, quiet=False):
    """Copy every .py / .yaml under `src` to `dest` (same relative paths) and write MANIFEST.json with their
    sha256.  Returns the number of files staged; 0 (and nothing touched) when `src` does not exist - the GPU box uses
    the files staged in the build container."""
    if not os.path.isdir(src):
        return 0
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    manifest = {}
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "old_maps", "maps_bin")]
        for f in sorted(files):
            if not f.endswith(KEEP_EXT):
                continue
            sp = os.path.join(root, f)
            rel = os.path.relpath(sp, src)
            dp = os.path.join(dest, rel)
            os.makedirs(os.path.dirname(dp), exist_ok=True)
            shutil.copyfile(sp, dp)
            with open(sp, "rb") as fh:
                manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(os.path.dirname(dest), "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    if not quiet:
        print("staged %d reference files from %s into %s" % (len(manifest), src, dest))
    return len(manifest)


if __name__ == "__main__":
    src = DEFAULT_SRC
    if "--src" in sys.argv:
        src = sys.argv[sys.argv.index("--src") + 1]
    n = stage(src)
    if n == 0:
        print("reference tree %s not found: nothing staged" % src)
