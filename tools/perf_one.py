"""run one probe config by name: python tools/perf_one.py <name> [size] [steps]"""
import os, sys
sys.argv = [sys.argv[0], sys.argv[2] if len(sys.argv) > 2 else "512", sys.argv[3] if len(sys.argv) > 3 else "3", "none"] + sys.argv[1:2]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "perf_probe.py")).read())
name = sys.argv[4]
if name == "dense":
    run("dense fast", 19, lb.EQ_FP32, (Z, Z, Z), BB, False, False)
elif name == "tasks_plain":
    run("tasks, source only", 19, lb.EQ_FP32, (Z, Z, Z), BB, False, True)
elif name == "tasks_obst":
    run("tasks, obstacle", 19, lb.EQ_FP32, (Z, Z, Z), BB, True, False)
elif name == "headline":
    run("headline", 19, lb.EQ_FP32, (Z, W, W), BB, True, True)
elif name == "inplace":
    run("in-place periodic", 19, lb.EQ_FP32, (P, P, P), BB, False, False, engine=lb.ENGINE_INPLACE)
elif name == "inplace_walls":
    run("in-place walls", 19, lb.EQ_FP32, (W, W, W), BB, False, False, engine=lb.ENGINE_INPLACE)
elif name == "walls_yz":
    run("general, walls y/z", 19, lb.EQ_FP32, (Z, W, W), BB, False, False)
elif name == "walls_y":
    run("general, walls y", 19, lb.EQ_FP32, (Z, W, Z), BB, False, False)
elif name == "periodic":
    run("general, periodic xyz", 19, lb.EQ_FP32, (P, P, P), BB, False, False)
elif name == "inplace_obst":
    run("in-place obstacle+walls+source", 19, lb.EQ_FP32, (P, W, W), BB, True, True, engine=lb.ENGINE_INPLACE)
