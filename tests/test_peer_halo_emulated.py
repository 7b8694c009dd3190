"""Peer-memory halos below the C ABI (csrc/lbm_halo.cu) on the host build of the library (tests/_emu.py): several slabs
in one process, connected with lbm_halo_export / lbm_halo_connect and stepped with lbm_group_step, must give the
single-domain answer -- bit for bit against one handle of the same library, and within the usual tolerance against the
CPU oracle.  The same bodies run on a GPU in tests/test_gpu_peer_halo.py."""
import numpy as np
import pytest

import _emu
import _oracle as O
import _peer_halo_cases as cases


@pytest.fixture(scope="module")
def lib():
    with _emu.emulated_library() as lib:
        yield lib


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_slab_group_equals_single_domain(lib, name):
    cases.run_case(name, devices=lambda world: [0] * world)


@pytest.mark.parametrize("world,fused", [(2, True), (3, True), (4, False)])
def test_reference_mode_on_slabs(lib, world, fused):
    cases.run_reference_mode_case(world, devices=lambda w: [0] * w, steps=8 if fused else 4, fused=fused)


def test_connect_refuses_mismatched_slabs(lib):
    cases.check_refusals()
