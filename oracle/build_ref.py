"""TEST INFRASTRUCTURE ONLY.  Recipe for ``oracle/_ref/``: the two reference modules that hold the hot path
(``midastouch/modules/particle_filter.py`` and ``midastouch/modules/pose.py``) are copied VERBATIM from the
read-only reference tree into ``oracle/_ref/midastouch/modules/`` so that the unmodified reference functions can be
timed on the GPU box's host cores (``bench.py --impl reference``, ``cpu_baseline.kind = "reference"``) -- the
reference tree itself does not exist there.  ``oracle/_ref/`` is git-ignored (no reference source enters the
history) but travels with the gpurun snapshot like a built ``.so``.

    python -m oracle.build_ref            # no-op when /root/reference is absent

The copies are loaded through ``oracle/ref_shim.py`` (stub modules for the absent trimesh / theseus / omegaconf)."""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("MIDAS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
# particle_filter.py / pose.py: the functions the CPU arm times.  filter.py: its loop statements (filter.py:150-190) are
# extracted as text and executed verbatim on the drop-in classes by tests/test_gpu_reference_loop.py.
FILES = ("midastouch/modules/particle_filter.py", "midastouch/modules/pose.py", "midastouch/filter/filter.py")


def main() -> int:
    if not all(os.path.isfile(os.path.join(SRC, f)) for f in FILES):
        print("oracle/build_ref: reference tree not present at %s (nothing to do)" % SRC)
        return 0
    sums = []
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), dst)
        sums.append("%s  %s" % (hashlib.sha256(open(dst, "rb").read()).hexdigest(), f))
    with open(os.path.join(DST, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(sums) + "\n")
    print("oracle/build_ref: copied %d reference modules into %s" % (len(FILES), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
