import sys, os
sys.path.insert(0, "tests")
import numpy as np, _oracle as O, _ref_gpu as R
n=16; spacing=float(np.float32(1.0/n)); path=os.path.join(O.GOLDEN_DIR,"obj","cube_angled.obj")
flow=np.full(27,0.001,np.float32); flow[6]=0.002
ref=R.RefLattice(n,n,n,tau=1.0,spacing=spacing)
cells=O.new_cells(n,n,n); cells["f"]=flow[None,:]
ref.load(cells); ref.insert_model(path); ref.preprocess()
mode=sys.argv[1]
if "flow" in mode: ref.set_flow(2, flow)
if "force" in mode: ref.set_reflection_data()
ref.prime(); ref.run_cli(3)
print(mode, "done", R.lib().ref_last_cuda_error())
