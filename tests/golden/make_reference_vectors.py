"""Regenerates tests/golden/reference_vectors.npz: outputs of the REFERENCE ITSELF.

The reference's own Lattice class and CUDA kernels (unmodified sources under /root/reference) are compiled for the
host on the SIMT executor of tests/emu/ (tests/_emu.py reference_library(), output oracle/_ref/libref_simt.so) and
run here on seeded inputs; the results are committed as small fixtures so that the oracle and the CUDA kernels can be
checked against the reference on machines that have neither /root/reference nor the prebuilt oracle/_ref.
Arithmetic note: g++ builds these kernels without FMA contraction (nvcc contracts), so the collision differs from a
GPU run of the reference in the last bit or two; streaming, masks and tables are exact.

    python tests/golden/make_reference_vectors.py        # needs /root/reference
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import _emu          # noqa: E402
import _oracle as O  # noqa: E402
import _ref_gpu as R  # noqa: E402


def main():
    so = _emu.reference_library()
    if so is None or not os.path.isdir("/root/reference/simulation/source"):
        raise SystemExit("needs /root/reference")
    R.SO, R._lib = so, None
    out = {}
    # 1. one streaming step and one collision (tau 1.0 and 0.6) on a seeded 8 x 6 x 8 lattice
    X, Y, Z = 8, 6, 8
    rng = np.random.default_rng(20261017)
    cells = O.new_cells(X, Y, Z)
    w = O.weights(27).astype(np.float32)
    cells["f"] = (w[None, :] * (1.0 + 0.1 * rng.standard_normal((X * Y * Z, 27)))).astype(np.float32)
    out["small_shape"] = np.array([X, Y, Z], np.int32)
    out["small_f0"] = cells["f"].copy()
    ref = R.RefLattice(X, Y, Z)
    ref.load(cells); ref.prime(); ref.stream()
    out["small_streamed"] = ref.retrieve()["f"].copy()
    ref.close()
    for tau in (1.0, 0.6):
        ref = R.RefLattice(X, Y, Z, tau=tau)
        ref.load(cells); ref.prime(); ref.collide()
        out["small_collided_tau%02d" % int(tau * 10)] = ref.retrieve()["f"].copy()
        ref.close()
    # 2. insertModel + preProcessModel of the tapered hexahedron at 12^3, then the CLI loop for 20 steps
    n, tau, steps = 12, 0.8, 20
    spacing = float(np.float32(1.0 / n))
    path = os.path.join(O.GOLDEN_DIR, "obj", "cube_angled.obj")
    flow = np.full(27, 0.001, np.float32); flow[6] = 0.002
    ref = R.RefLattice(n, n, n, tau=tau, spacing=spacing)
    start = O.new_cells(n, n, n)
    start["f"] = flow[None, :]
    ref.load(start)
    ref.insert_model(path)
    ref.preprocess()
    ref.set_flow(2, flow)
    ref.prime()
    first = ref.retrieve()
    tables, cells_with_table = ref.tables(first)
    ref.run_cli(steps)
    last = ref.retrieve()
    ref.close()
    out["cli_n_tau_steps"] = np.array([n, tau, steps], np.float64)
    out["cli_flow"] = flow
    out["cli_is_internal"] = first["is_internal"].copy()
    out["cli_is_reflected"] = (first["is_reflected"] != 0).astype(np.uint8)
    out["cli_table_cells"] = cells_with_table.astype(np.int64)
    out["cli_table_dirs"] = tables["dirs"].copy()
    out["cli_table_w"] = tables["w"].copy()
    out["cli_f_final"] = last["f"].copy()
    dst = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(dst, **out)
    print(dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
