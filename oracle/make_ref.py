"""Stage the UNMODIFIED reference sources of the hot path under oracle/_ref/ (build container only).

    python oracle/make_ref.py

TEST INFRASTRUCTURE ONLY.  The GPU box has no /root/reference, so the three files the path consists of —
``reward.py`` (compute_orie / compute_dcsb), ``lib/metrics.py`` (box_correct, ap_per_class, compute_ap) and
``lib/data.py`` (load_data / set_data) — are copied verbatim, where they lie, into ``oracle/_ref/`` so that
``bench.py --impl reference`` and ``cpu_baseline`` can time the reference ITSELF on the box's host cores.
``oracle/_ref/`` is listed in .gitignore (the sources never enter this repository's history) but not in
.gpurunignore (it travels to the GPU box like the built libraries).  Nothing in the product package imports it;
``oracle/ref_local.py`` is the only loader.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("ORIE_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ["reward.py", os.path.join("lib", "__init__.py"), os.path.join("lib", "metrics.py"), os.path.join("lib", "data.py")]


def main() -> int:
    if not os.path.isfile(os.path.join(REF_ROOT, "reward.py")):
        print(f"make_ref: {REF_ROOT} not present; leaving {OUT} as it is")
        return 0
    os.makedirs(os.path.join(OUT, "lib"), exist_ok=True)
    lines = []
    for f in FILES:
        src, dst = os.path.join(REF_ROOT, f), os.path.join(OUT, f)
        shutil.copyfile(src, dst)
        lines.append(f"{hashlib.sha256(open(dst, 'rb').read()).hexdigest()}  {f}")
    with open(os.path.join(OUT, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print(f"make_ref: staged {len(FILES)} files from {REF_ROOT} into {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
