"""The far side of the `call` seam: himut_b200.vcfio.dump_sbs / dump_phased_sbs must write, byte for byte, what the
reference's vcflib.dump_sbs / dump_phased_sbs write from the same rows (tests/golden/vcf_text.json holds the digests
of the reference's output for the rows of every committed `call_*` fixture)."""
import importlib.util
import json
import os

import pytest

import cases
import refshim
from himut_b200 import vcfio

_spec = importlib.util.spec_from_file_location("mkvcf", os.path.join(cases.GOLDEN_DIR, "make_golden_vcf.py"))
mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mk)
EXPECTED = json.load(open(os.path.join(cases.GOLDEN_DIR, "vcf_text.json")))["expected"]


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_writer_text_matches_the_reference(name):
    rows = mk.rows_of(name)
    assert mk.digest(vcfio.dump_sbs, rows) == EXPECTED[name]["unphased"]
    assert mk.digest(vcfio.dump_phased_sbs, rows) == EXPECTED[name]["phased"]


def test_fixtures_cover_the_quirky_rows():
    statuses = {r[4] for name in EXPECTED for r in mk.rows_of(name)}
    assert {"PASS", "HetAltSite", "LowGQ"} <= statuses
    assert any(r[11] != "." for r in mk.rows_of("call_phase"))


def test_suffix_is_checked(tmp_path):
    with pytest.raises(ValueError):
        vcfio.dump_sbs(str(tmp_path / "out.txt"), mk.HEADER, [], {})


@pytest.mark.skipif(not refshim.have_reference(), reason="reference sources are not present")
def test_digests_are_the_reference_writers_output():
    refshim.import_reference()
    import himut.vcflib as ref
    for name in ("call_adversarial_a", "call_phase"):
        rows = mk.rows_of(name)
        assert mk.digest(ref.dump_sbs, rows) == EXPECTED[name]["unphased"]
        assert mk.digest(ref.dump_phased_sbs, rows) == EXPECTED[name]["phased"]
