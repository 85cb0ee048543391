#!/usr/bin/env python
"""Test infrastructure: runs the reference's own command line (himut.__main__.main, /root/reference/src) in this
process, either untouched ("reference") or with himut_b200.patch.install() ("dropin"), on the data set of
cases.cli_dataset().  Build container only — the reference is not on the GPU box.

    python tests/cli_runner.py reference|dropin <workdir> <phase_block or 0> <himut argv ...>

Both modes read the BAM path through the pysam look-alike of tests/shims (the reference's driver code opens it for the
header and, in "reference" mode, for everything else); in "dropin" mode the workers read the real BAM file with the
native decoder.  Without a GPU the drop-in workers' device calls are answered by tests/standin.OracleContext; with
HIMUT_B200_CLI_REAL_CONTEXT=1 they go to the CUDA library.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    mode, workdir, phase_block = sys.argv[1], sys.argv[2], int(sys.argv[3])
    argv = sys.argv[4:]
    import __graft_entry__ as g
    g.build()
    import cases
    import refshim
    himut = refshim.import_reference()
    import pysam
    bam = argv[argv.index("-i" if "-i" in argv else "--bam") + 1]
    data = cases.cli_dataset(phase_block or None)
    pysam.register(bam, refshim.ContigsProvider([(c, n, d.batch) for c, n, d in data]))
    himut.gtlib.np = refshim.StableArgsortNumpy()
    from himut_b200 import patch
    if mode == "dropin":
        patch.install()
        if os.environ.get("HIMUT_B200_CLI_REAL_CONTEXT") != "1":
            import standin
            from himut_b200 import worker
            ctx = standin.OracleContext()
            worker.context = lambda: ctx  # inherited by the pool workers (fork)
            plain = worker.RegionSource.batch  # the oracle reads the bases: ask the decoder for them
            worker.RegionSource.batch = lambda self, chrom, loci, phase_sets=None, seq=True, **kw: plain(self, chrom, loci, phase_sets, seq=True, **kw)
    else:
        patch._scipy_compat()  # himut.__main__ imports phaselib, which needs scipy.stats.binom_test (removed in scipy 1.12)
    os.chdir(workdir)  # himut.log / norm.log are written to the working directory
    sys.argv = ["himut"] + argv
    import himut.__main__
    himut.__main__.main()


if __name__ == "__main__":
    main()
