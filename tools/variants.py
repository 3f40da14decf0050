#!/usr/bin/env python
"""Build experiment variants of the library next to the default one and (on a GPU box) compare them.

    python tools/variants.py build coop=-DIRBPP_COOP_APPROX fine=-DIRBPP_PROBE_FINE
        -> irbpp_b200/lib/libirbpp_coop.so, libirbpp_fine.so   (they travel with gpurun like the default .so)
    python tools/variants.py bench coop            # GPU: parity tests of the episode goldens + bench.py per variant

First GPU call of a round (everything below is checked under the CUDA emulator, none of it is timed yet):
    python tools/variants.py build coop=-DIRBPP_COOP_APPROX split=-DIRBPP_SPLIT_APPLY \\
        split10=-DIRBPP_SPLIT_APPLY,-DIRBPP_SCAN_MIN_CTAS=10 both=-DIRBPP_COOP_APPROX,-DIRBPP_SPLIT_APPLY \\
        pf=-DIRBPP_PREFETCH_NEXT=1184
    python tools/variants.py bench coop split split10 both pf

Known switches (csrc/): IRBPP_COOP_APPROX  contours of >= 17 points by a whole warp (dp_keep_warp)
                        IRBPP_PROBE_FINE   slots 4-7 of the phase counters time the sub-phases of phase C
                        IRBPP_SPLIT_APPLY  phase A (apply the action) as its own one-warp-per-bin kernel in front of the scan
                        IRBPP_PREFETCH_NEXT=1184   scan CTAs prefetch the inputs of the bin a later wave handles into L2
                        IRBPP_ENVS_PER_CTA / IRBPP_CAND_WARPS / IRBPP_SCAN_MIN_CTAS   CTA shapes
Run-time knobs of the host path (environment, read at irbpp_create; defaults = what bench.py's e2e measures):
    IRBPP_HOST_RESULTS=stores|kernel|memcpy   per-bin PCIe stores (default) | one coalesced copy kernel | D2H memcpy in step_wait
    IRBPP_HOST_ACTIONS=mapped|memcpy          kernel reads the pinned actions over PCIe (default) | H2D memcpy in front of it
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB_DIR = os.path.join(ROOT, "irbpp_b200", "lib")


def lib_path(name):
    return os.path.join(LIB_DIR, "libirbpp.so" if name == "default" else "libirbpp_%s.so" % name)


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    cmd, args = sys.argv[1], sys.argv[2:]
    if cmd == "build":
        from irbpp_b200 import build as b
        for spec in args:
            name, defs = spec.split("=", 1)
            print(b.build(out=lib_path(name), defs=defs.split(",")))
        return 0
    if cmd == "bench":
        for name in ["default"] + args:
            env = dict(os.environ, IRBPP_LIB=lib_path(name))
            t = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x",
                                "-m", "gpu", "-k", "episode or random or hull"], env=env, capture_output=True, text=True)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "100", "--warmup", "5",
                                "--no-cpu-baseline"], env=env, capture_output=True, text=True)
            line = json.loads(r.stdout.strip().splitlines()[-1])
            print(json.dumps({"variant": name, "tests": t.stdout.strip().splitlines()[-1], "ms_per_step": line["ms_per_step"],
                              "e2e_ms": line["e2e"]["ms_per_step"]}))
        return 0
    raise SystemExit(__doc__)


if __name__ == "__main__":
    sys.exit(main())
