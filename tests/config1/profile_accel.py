"""cProfile of the accelerated config-1 pipeline: where the time of the bound hot-path functions
goes inside the reference's Pipelines.Analysis (test infrastructure, like run_config1.py)."""
import cProfile
import os
import pstats
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_config1 as R

work = os.path.abspath(sys.argv[1])
if not os.path.exists(os.path.join(work, "expression.txt")):
    R.main(["make", work])
pr = cProfile.Profile()
pr.enable()
R.run_pipeline(work, "accel")
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats("fastproject_b200|Signatures|_engine|_scoring", 40)
