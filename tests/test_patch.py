"""himut_b200.patch.install() against the unmodified reference package (build container only: the reference is
not on the GPU box): every seam named in INTEGRATION.md is re-pointed, with matching signatures."""
import inspect

import pytest

import refshim


@pytest.mark.skipif(not refshim.have_reference(), reason="reference sources are not present")
def test_install_repoints_every_seam():
    himut = refshim.import_reference()
    import himut.bamlib
    import himut.caller
    import himut.normcounts
    import himut.reflib
    originals = {
        "caller.get_somatic_substitutions": himut.caller.get_somatic_substitutions,
        "normcounts.get_callable_tricounts": himut.normcounts.get_callable_tricounts,
        "reflib.get_chrom_tricount": himut.reflib.get_chrom_tricount,
        "bamlib.get_thresholds": himut.bamlib.get_thresholds,
    }
    from himut_b200 import patch
    patch.install()
    import himut.phaselib  # importable only after the scipy compatibility shim
    import himut_b200.bamlib, himut_b200.caller, himut_b200.normcounts, himut_b200.phaselib, himut_b200.reflib
    assert himut.caller.get_somatic_substitutions is himut_b200.caller.get_somatic_substitutions
    assert himut.normcounts.get_callable_tricounts is himut_b200.normcounts.get_callable_tricounts
    assert himut.reflib.get_chrom_tricount is himut_b200.reflib.get_chrom_tricount
    assert himut.bamlib.get_thresholds is himut_b200.bamlib.get_thresholds
    assert himut.phaselib.get_edges is himut_b200.phaselib.get_edges
    # same positional parameters as the functions they replace
    for name, orig in originals.items():
        mod, fn = name.split(".")
        ours = getattr(getattr(himut, mod), fn)
        assert list(inspect.signature(ours).parameters) == list(inspect.signature(orig).parameters), name
    # restore, so later tests in this process see the reference's own workers again
    himut.caller.get_somatic_substitutions = originals["caller.get_somatic_substitutions"]
    himut.normcounts.get_callable_tricounts = originals["normcounts.get_callable_tricounts"]
    himut.reflib.get_chrom_tricount = originals["reflib.get_chrom_tricount"]
    himut.bamlib.get_thresholds = originals["bamlib.get_thresholds"]
