"""Re-points the reference's starmap targets (and its BAM pre-pass) at the GPU workers.

The reference resolves its workers as module globals at call time (src/himut/caller.py:805-806,
src/himut/normcounts.py:536), so assigning the attributes before the driver runs is enough; the
CLI, header, thresholds, natsort and writers stay the reference's own code.
"""


def _scipy_compat():
    """scipy >= 1.12 dropped scipy.stats.binom_test, which himut.phaselib imports at module level
    (src/himut/phaselib.py:11); without it `himut.__main__` cannot be imported.  Same test, new name."""
    import scipy.stats
    if not hasattr(scipy.stats, "binom_test"):
        def binom_test(x, n=None, p=0.5, alternative="two-sided"):
            return scipy.stats.binomtest(int(x), int(n), p, alternative=alternative).pvalue
        scipy.stats.binom_test = binom_test


def install():
    _scipy_compat()
    import himut.bamlib
    import himut.caller
    import himut.normcounts
    import himut.phaselib
    import himut.reflib

    from . import bamlib, caller, normcounts, phaselib, reflib
    himut.caller.get_somatic_substitutions = caller.get_somatic_substitutions
    himut.normcounts.get_callable_tricounts = normcounts.get_callable_tricounts
    himut.reflib.get_chrom_tricount = reflib.get_chrom_tricount  # reflib.py:42-52 starmap target
    himut.bamlib.get_thresholds = bamlib.get_thresholds          # BAM pre-pass, called at caller.py:690
    himut.phaselib.get_edges = phaselib.get_edges                # `himut phase` pair tables, called at phaselib.py:248
    return himut


def main():
    """console entry: `python -m himut_b200.patch call -i in.bam ...` = `himut call ...` on GPUs"""
    install()
    import himut.__main__
    himut.__main__.main()


if __name__ == "__main__":
    main()
