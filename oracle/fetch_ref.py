"""Recipe that puts the UNTOUCHED reference next to the oracle so that it travels to the GPU box (TEST INFRASTRUCTURE).

The reference (agent-charon/MinGraph-UNet) is pure Python: there is nothing to compile and no ``setup.py`` to pip-install,
so "building" it is a verbatim copy of its source tree from ``/root/reference/MinGraph-UNet`` into the git-ignored
``oracle/_ref/MinGraph-UNet`` (never committed: ``.gitignore`` lists ``oracle/_ref/``; ``gpurun`` snapshots it like the
built ``.so`` files).  ``__graft_entry__.build()`` runs this in the build container, where ``/root/reference`` exists;
on the GPU box only the copy is there.  ``oracle/ref_loader.py`` imports the classes from it IN PLACE — nothing in
the product package reads it; ``bench.py --impl reference`` times it (``cpu_baseline.kind = "reference"``) and the
live-reference tests of ``tests/test_oracle.py`` / ``tests/test_host.py`` run against it.

    python oracle/fetch_ref.py [SRC]      # SRC defaults to /root/reference/MinGraph-UNet
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SRC = "/root/reference/MinGraph-UNet"
DST = os.path.join(HERE, "_ref", "MinGraph-UNet")
KEEP_EXT = (".py", ".yaml", ".yml", ".txt", ".md")


def fetch(src: str = DEFAULT_SRC, dst: str = DST) -> int:
    """Copies the source files (by extension; no data, no caches).  Returns the number of files, 0 when ``src`` is absent."""
    if not os.path.isfile(os.path.join(src, "model", "gat", "graph_attention.py")):
        return 0
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", ".git")]
        rel = os.path.relpath(root, src)
        for f in files:
            if f.endswith(KEEP_EXT):
                os.makedirs(os.path.join(dst, rel), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(dst, rel, f))
                n += 1
    return n


if __name__ == "__main__":
    n = fetch(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_SRC)
    print(f"oracle/_ref: {n} files" if n else "reference tree not found; oracle/_ref left as it is")
