"""Stage the reference's Python sources for a GPU run WITHOUT adding them to this repository.

    python tools/stage_reference.py            (build container only: /root/reference must exist)

Writes oracle/_ref/reference_src.tar.gz (git-ignored, not gpurun-ignored, so it travels to the GPU box with the snapshot
like a built .so) holding the reference's src/ package and its tests/gp/*.py unit tests -- no data files.
tools/run_reference_modules.py and tests/test_gpu_reference_modules.py unpack it into the temp directory of the GPU box and
run the reference's own modules, unmodified, on the import shim."""
import os
import sys
import tarfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("BATTGP_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref", "reference_src.tar.gz")


def main():
    if not os.path.isdir(os.path.join(REF, "src")):
        print(f"{REF}/src not found", file=sys.stderr)
        return 1
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    n = 0
    with tarfile.open(OUT, "w:gz") as tf:
        for top in ("src", os.path.join("tests", "gp")):
            for dirpath, dirs, files in os.walk(os.path.join(REF, top)):
                dirs[:] = [d for d in dirs if d != "__pycache__"]
                for f in files:
                    if f.endswith(".py"):
                        p = os.path.join(dirpath, f)
                        tf.add(p, arcname=os.path.relpath(p, REF))
                        n += 1
        init = os.path.join(REF, "tests", "__init__.py")
        if os.path.exists(init):
            tf.add(init, arcname="tests/__init__.py")
    print(f"{OUT}: {n} files, {os.path.getsize(OUT)} bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
