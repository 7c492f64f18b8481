"""TEST INFRASTRUCTURE — recipe that makes the UNMODIFIED reference runnable on the GPU box.

The reference (tan9zj/xnrs) is pure Python, so "compiling it where it lies" is a file copy: this script copies the
package `/root/reference/xnrs` byte for byte into `oracle/_ref/xnrs` (git-ignored, NOT gpurun-ignored: it travels to the
GPU box with the snapshot like the built .so files, and never enters the repo's history).  `__graft_entry__.build()` runs
it whenever `/root/reference` is present; on the GPU box only the copy is used.  Consumers: `bench.py --impl reference`
and bench.py's `cpu_baseline` legs (kind "reference"), through `oracle/refload.py` — nothing under `xnrs_b200/` may touch it.

    python oracle/make_ref.py            # -> oracle/_ref/xnrs/** and oracle/_ref/MANIFEST.json (sha256 per file)
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/xnrs'
DST = os.path.join(HERE, '_ref')


def make(verbose: bool = False) -> bool:
    """-> True if the copy exists afterwards (False: no reference here and no earlier copy)"""
    if not os.path.isdir(SRC):
        return os.path.isdir(os.path.join(DST, 'xnrs'))
    out = os.path.join(DST, 'xnrs')
    if os.path.isdir(out):
        shutil.rmtree(out)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(SRC, out, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    manifest = {}
    for dirpath, _, files in os.walk(out):
        for f in sorted(files):
            p = os.path.join(dirpath, f)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, 'rb').read()).hexdigest()
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as fh:
        json.dump({'source': SRC, 'files': manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f'oracle/_ref: {len(manifest)} reference files copied unmodified from {SRC}')
    return True


if __name__ == '__main__':
    sys.exit(0 if make(verbose=True) else 1)
