"""Oracle + generator against the committed digests (tests/golden/synthetic_digests.json)."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
CASES = json.load(open(os.path.join(HERE, "golden", "synthetic_digests.json")))


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"shape{c['shape']}-{c['records']}-l{c['level']}")
def test_oracle_matches_committed_digest(case):
    from make_golden import digest_case
    got = digest_case(case["shape"], case["records"], case["level"], case["gc_seed"])
    assert got == case
