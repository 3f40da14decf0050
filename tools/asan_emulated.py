#!/usr/bin/env python
"""AddressSanitizer over the kernels under the CUDA emulator: a CPU-side memcheck of every shared / global
access (dynamic shared memory is a heap block of exactly the launch's size, so a wrong size shows up).

    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 \\
        python tools/asan_emulated.py [-DIRBPP_TASKS_PER_LANE=2 ...] 2>&1 | grep -E "ERROR|SUMMARY| ok"
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import test_kernels_emulated as t  # noqa: E402

emu = t.build_emulated(tempfile.mkdtemp(), defs=sys.argv[1:] + ["-fsanitize=address", "-g", "-fno-omit-frame-pointer"])
t.test_emulated_hull_actions_match_reference_golden(emu); print("hull goldens ok")
t.test_emulated_survey_known_answers(emu); print("known answers ok")
t.test_emulated_many_start_pixels(emu); print("many start pixels ok")
t.test_emulated_heuristics_match_reference_golden(emu); print("heuristics ok")
for tag in ("blockout", "irregular", "cube"):
    t.test_emulated_scan_matches_reference_golden(emu, tag); print("scan", tag, "ok")
for name, steps in (("episode_blockout", 12), ("episode_irregular", 4), ("episode_truncate", 4), ("episode_buffered", 5),
                    ("episode_rot24", 2)):
    t._replay(emu, name, steps); print(name, "ok")
