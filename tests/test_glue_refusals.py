"""The reference-side binding (oracle/ref_glue.cpp) hands a Domain it cannot reproduce back to the CPU
integrator instead of computing something else: these refusals happen while the model is read out of the Domain, before
any device call, so they run without a GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
from modelspec import GLUE_SO, RefBackend, frame2d, have_glue, have_ref, soil_column_equaldof, with_corot, with_joint_offsets, J2_STEEL  # noqa: E402

pytestmark = pytest.mark.skipif(not (have_ref() and have_glue()), reason="oracle/_ref not built (needs /root/reference)")


def _setup(spec, **kw):
    D = RefBackend(spec, defer_setup=True, so=GLUE_SO, **kw)
    D.setup_glue_loadcontrol(1, 0, 0.125)
    return D


def test_corotational_with_joint_offsets_is_refused():
    with pytest.raises(RuntimeError, match="joint offsets"):
        _setup(with_joint_offsets(with_corot(frame2d(1, 1, 1)), seed=1))
