import sys, json
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
print(json.dumps(bench.tree_builder_leg(0, sizes=((1000, 3),)), indent=1))
