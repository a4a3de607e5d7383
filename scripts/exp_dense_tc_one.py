"""a few full-size forward + dX launches for ncu: python scripts/exp_dense_tc_one.py <act> <p>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scripts.exp_dense_tc as e
act, p = int(sys.argv[1]), float(sys.argv[2])
e.run(29858, 40981, 8, act, p)
