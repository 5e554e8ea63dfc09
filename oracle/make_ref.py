"""Recipe for oracle/_ref/: an UNMODIFIED copy of the reference program, made from the read-only checkout in the build
container so that it travels to the GPU box with the repository snapshot (oracle/_ref/ is git-ignored: reference sources never
enter this repository's history; it is not gpurun-ignored).

The reference (tsnuk/trade-AId-multimodal-transformer) is pure Python: there is nothing to compile, "building" it is copying
its modules, its two YAML files and its demo (examples/). Used as the checker only:
  * tests/test_main_py_gpu.py runs its main.py on the GPU box unmodified and with the INTEGRATION.md shims;
  * tests/test_generate_gpu.py compares `generate` token by token with its model.py;
  * bench.py --impl reference / --impl reference-eager time its own model.py + training_utils.py (cpu_baseline.kind "reference").
Nothing of the product imports it.

usage: python oracle/make_ref.py [/root/reference]"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
KEEP_EXT = (".py", ".yaml", ".yml", ".csv", ".md", ".txt")


def make(src="/root/reference", dest=DEST):
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src}: the reference checkout exists in the build container only")
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    n = 0
    for d, dirs, files in os.walk(src):
        dirs[:] = [x for x in dirs if x not in (".git", "__pycache__", "tests")]
        rel = os.path.relpath(d, src)
        out = os.path.join(dest, rel) if rel != "." else dest
        os.makedirs(out, exist_ok=True)
        for f in files:
            if f.endswith(KEEP_EXT) or f == "LICENSE":
                shutil.copyfile(os.path.join(d, f), os.path.join(out, f))
                os.chmod(os.path.join(out, f), 0o644)
                n += 1
    with open(os.path.join(dest, "PROVENANCE.txt"), "w") as f:
        f.write(f"unmodified copy of {src} made by oracle/make_ref.py; {n} files\n")
    return dest


def available():
    return os.path.isfile(os.path.join(DEST, "model.py")) and os.path.isfile(os.path.join(DEST, "main.py"))


if __name__ == "__main__":
    print(make(*(sys.argv[1:2])))
