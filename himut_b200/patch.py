"""Re-points the reference's starmap targets (and its BAM pre-pass) at the GPU workers.

The reference resolves its workers as module globals at call time (src/himut/caller.py:805-806,
src/himut/normcounts.py:536), so assigning the attributes before the driver runs is enough; the
CLI, header, thresholds, natsort and writers stay the reference's own code.
"""


def install():
    import himut.caller
    import himut.normcounts
    import himut.bamlib
    import himut.reflib

    from . import bamlib, caller, normcounts, reflib
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
