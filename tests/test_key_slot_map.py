"""CPU test of the host key map (pathway_b200/csrc/key_slot_map.h = KeyToU64IdMapper, mod.rs:209-256): a g++-built
differential test against std::unordered_map plus the churn-round micro-benchmark."""

import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_key_slot_map_differential(tmp_path):
    exe = str(tmp_path / "ksm_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "native", "key_slot_map_test.cpp"), "-o", exe])
    out = subprocess.run([exe, "200000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr


def test_batched_swap_remove_differential(tmp_path):
    """pathway_b200/csrc/slot_book.h (parallel hash-table phases + sequential swap replay) against the literal loop of
    brute_force_knn_integration.rs:146-170: statuses, final slot -> key layout, key map and the row moves."""
    exe = str(tmp_path / "slot_book_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-pthread",
                           os.path.join(ROOT, "tests", "native", "slot_book_test.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr


def test_host_pool_regions(tmp_path):
    """pathway_b200/csrc/host_pool.h: every part of every region runs exactly once (back-to-back regions with polling
    workers, and regions after an idle gap with sleeping workers)."""
    exe = str(tmp_path / "host_pool_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-pthread",
                           os.path.join(ROOT, "tests", "native", "host_pool_test.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
