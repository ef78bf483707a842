"""Build container only: the oracle against the unmodified reference imported from /root/reference
(bit-identical float64 output on identical features and identical noise buffers)."""
import tempfile

import numpy as np
import pytest

from oracle import ref_harness
from oracle import validate_against_reference as V

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_harness.reference_available(), reason="/root/reference is not present")]


@pytest.mark.parametrize("case", [V.CASES[0], V.CASES[2], V.CASES[8], V.CASES[10]], ids=lambda c: c[0])
def test_oracle_is_bit_identical_to_reference(case):
    name, si, secs, cli = case
    with tempfile.TemporaryDirectory() as tmp:
        ref, orc = V.run_case(tmp, name, si, secs, cli)
    assert ref.shape == orc.shape and ref.dtype == np.float64
    assert np.array_equal(ref, orc)
