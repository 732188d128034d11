"""The driver's smoke entry point must stay runnable: one tiny rollout + update on cuda:0, checked against the oracle."""
import pytest

pytestmark = pytest.mark.gpu


def test_graft_entry_smoke_runs():
    import __graft_entry__ as entry
    entry.smoke()
