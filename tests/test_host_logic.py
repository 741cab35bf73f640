"""CPU suite: the host logic of multiprime_b200.core driven through tests/fake_device.py (no GPU), compared with
the reference's golden records window by window."""
import pytest

from tests import fake_device
from tests.parity import check_case


@pytest.mark.parametrize("name,limit", [("synth_iupac", 60), ("synth300", 40), ("c2_k18", 12), ("c3_tmsa", 14),
                                        ("c1_testfa", 45)])
def test_host_logic_fake_device(name, limit):
    check_case(name, backend=fake_device, max_windows=limit)
