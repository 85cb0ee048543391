"""One rank of a sharded `himut call` over a multi-contig BAM with the CUDA library (torchrun launches it):
rank 0 writes per-contig row digests, log vectors and the job statistics as JSON.  Used by tests/test_zz_gpu_genome.py."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    bam, out, contigs_json = sys.argv[1], sys.argv[2], sys.argv[3]
    import torch
    import torch.distributed as dist
    import parity
    from himut_b200 import genome, gtmodel
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    contigs = json.load(open(contigs_json))
    args = dict(gtmodel.DEFAULT_CALL_ARGS, non_human_sample=True)
    loci = {c: genome.chunkloci(c, n) for c, n in contigs}
    lst, log, stats = genome.call_genome(bam, loci, args)
    if lst is not None:
        json.dump({"digest": {c: parity.rows_digest(lst[c]) for c in lst}, "rows": {c: len(lst[c]) for c in lst},
                   "log": log, "stats": stats}, open(out, "w"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
