"""cProfile of the host side of two denoise steps (where do 0.2 s of enqueue time per step go?)."""
import cProfile
import pstats
import sys
import os
sys.argv = [sys.argv[0], "--steps", "2", "--warmup", "1"]
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import profile_step  # noqa: E402
pr = cProfile.Profile()
pr.enable()
profile_step.main()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
