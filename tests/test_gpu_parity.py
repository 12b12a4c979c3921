"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on identical
seeded inputs.  Bar: bit-exact float32 state, identical contact lists (pairs, order, touching,
point counts), identical stored impulses, fat AABBs, controller and light state, at every env-step.
"""
import numpy as np
import pytest

from gym_kilobots_b200 import _abi as abi
from gym_kilobots_b200 import scenarios as SC

from parity_util import run_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("toi", [False, True])
def test_c1_single_env(oracle, native, toi):
    run_parity(oracle, native, SC.c1_single_env(8, enable_toi=toi), steps=40)


@pytest.mark.parametrize("toi", [False, True])
def test_c2_quad_assembly_degenerate(oracle, native, toi):
    run_parity(oracle, native, SC.c2_quad_assembly(64, enable_toi=toi), steps=40)


def test_c2_prime(oracle, native):
    run_parity(oracle, native, SC.c2_quad_assembly(64, degenerate=False), steps=40)


def test_c3_shapes(oracle, native):
    run_parity(oracle, native, SC.c3_shapes(12, num_kilobots=40), steps=25)


def test_c5_small(oracle, native):
    run_parity(oracle, native, SC.c5_small(512), steps=40)
