"""TEST INFRASTRUCTURE: copies the reference's three analysis scripts, unmodified, from /root/reference into
tests/golden/_ref_scripts/ (git-ignored: no reference source enters this repository's history; the directory travels to the GPU
box with the working tree, like the built .so files).  tests/test_ref_scripts_gpu.py executes them with runpy against shims/.
Run here (the container that has /root/reference): `python tests/golden/make_ref_scripts.py`; __graft_entry__.build() calls it."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref_scripts")
FILES = {
    "debug_tda_pipeline.py": "debug_tda_pipeline.py",
    "analyze_tda_over_layers.py": "analyze_tda_over_layers.py",
    "experiments/adversarial_compositional_binding/analyze_adversarial_tda.py": "experiments/adversarial_compositional_binding/analyze_adversarial_tda.py",
}


def main(ref="/root/reference"):
    if not os.path.isdir(ref):
        print(f"{ref} not present: nothing copied")
        return False
    for src, dst in FILES.items():
        d = os.path.join(OUT, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(ref, src), d)
    print(f"copied {len(FILES)} reference scripts to {OUT}")
    return True


if __name__ == "__main__":
    main(*sys.argv[1:])
