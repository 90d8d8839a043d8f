"""SASS evidence of the built library, per kernel: TMA bulk copies (UBLKCP), mbarrier ops (SYNCS), local-memory spills
(STL / LDL), fp64 pipe ops, and the target architecture.  Writes profiles/<tag>_sass.json + a short text excerpt.

    python tools/sass_excerpt.py r02
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_trading_env_b200 import _cabi  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
cuobjdump = "/usr/local/cuda/bin/cuobjdump"
elf = subprocess.run([cuobjdump, "-lelf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
sass = subprocess.run([cuobjdump, "-sass", _cabi.LIB_PATH], capture_output=True, text=True).stdout
demangle = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
parts = re.split(r"\n\s*Function : ", sass)
out = {"library": os.path.relpath(_cabi.LIB_PATH, ROOT), "build_id": _cabi.built_id(), "elf": elf.strip().splitlines(), "kernels": {}}
ops = ["UBLKCP", "SYNCS", "STL", "LDL", "DFMA", "DMUL", "DADD", "MUFU.RCP64H", "LDG", "STG", "ATOMG", "MEMBAR", "CALL", "BAR.SYNC", "ACQBULK", "UTMALDG", "HMMA", "UTCHMMA"]
excerpt = []
for name, body in zip(demangle, parts[1:]):
    lines = [ln for ln in body.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
    counts = {op: sum(1 for ln in lines if re.search(r"\b" + re.escape(op) + r"\b", ln)) for op in ops}
    short = re.sub(r"\(.*", "", name).replace("gte::", "")
    out["kernels"][short] = {"sass_instructions": len(lines), **{k: v for k, v in counts.items() if v}}
    if "obs_tma_coop_kernel<3, 2, 4, 32, false, true>" in short:            # the claimed-tiles form: what C5 / C4 launch
        excerpt = [short] + [ln.strip() for ln in lines if re.search(r"UBLKCP|SYNCS|UTMALDG", ln)][:24]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_sass.json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", f"{tag}_sass_excerpt.txt"), "w").write("\n".join(excerpt) + "\n")
for k, v in out["kernels"].items():
    if any(s in k for s in ("obs_tma_coop_kernel<3, 2, 4, 32", "obs_tma_coop_kernel<3, 2, 4, 16, true", "step_kernel")):
        print(k, v)
