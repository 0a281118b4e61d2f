"""Builds a tuning variant of libicp_b200.so with extra -D flags into build/variants/<name>.so (git-ignored, travels
to the GPU box).  Run a tool against it with ICPB_LIB=build/variants/<name>.so.
usage: python tools/build_variant.py name [-DICPB_SEARCH_THREADS=128 ...]"""
import importlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
bld = importlib.import_module("slam-envire_b200.build")
name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "build", "variants")
os.makedirs(out, exist_ok=True)
lib = os.path.join(out, name + ".so")
cmd = [bld.nvcc_path()] + bld.NVCC_FLAGS + flags + ["-o", lib] + [os.path.join(bld.CSRC, s) for s in bld.SOURCES] + ["-ldl"]
subprocess.check_call(cmd)
print(lib)
