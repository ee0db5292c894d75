"""Recipe for oracle/_ref/: the reference's own hot-path module, as a git-ignored build artefact.

    python oracle/build_ref.py            # in the build container, where /root/reference exists

The reference is pure Python (no build system, nothing to compile), so its "built" form is the module file itself:
`/root/reference/models.py` is copied VERBATIM to `oracle/_ref/models.py`.  `oracle/_ref/` is listed in .gitignore (no
reference source enters the history) but not in .gpurunignore, so it travels to the GPU box like the repo's own
libmar.so — `/root/reference` does not exist there.  Test / measurement infrastructure only:

* `bench.py --impl reference` and `bench.py`'s `cpu_baseline` time the LIVE reference classes on the box's host cores
  (`cpu_baseline.kind = "reference"`); without `oracle/_ref` they fall back to the oracle port (`kind = "port"`);
* `tools/ref_on_b200.py` runs the same unmodified modules on the B200 under torch eager (the survey's kernel bar).

Nothing under multimodalaggressionrecognition_b200/ imports it."""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ("models.py",)          # models.py:91-175, 225-295, 344-430, 480-558, 667-770, 823-928 are the hot path


def build() -> bool:
    if not os.path.isdir(SRC):
        return os.path.exists(os.path.join(DST, FILES[0]))
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as m:
        for f in FILES:
            with open(os.path.join(DST, f), "rb") as fh:
                m.write(f"{hashlib.sha256(fh.read()).hexdigest()}  {f}  (verbatim copy of {SRC}/{f})\n")
    return True


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref ready" if ok else "no /root/reference here and no oracle/_ref: the CPU arm will use the oracle port")
    sys.exit(0)
