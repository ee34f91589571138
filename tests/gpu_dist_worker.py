"""Worker of tests/test_dist_gpu.py (one process per GPU under torch.distributed.run, NCCL): the row-sharded
query_topk + rerank_topk of sprc_b200/retrieval.py with the CUDA model must reproduce, bit for bit, the rows the
same model returns from the whole index on one GPU (computed first, before the process group exists)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sprc_b200 import retrieval as RT  # noqa: E402
from sprc_b200 import synth  # noqa: E402
from sprc_b200.model import Blip2QformerCirRerank  # noqa: E402


def main(out_dir):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N, Q, T, k = 41, 10, 6, 12
    m = Blip2QformerCirRerank(vit_model="clip_L", device=dev, max_images=64, max_queries=16, max_pairs=4 * T,
                              vit_depth=2, qf_layers=2)
    sd = synth.make_state_dict("clip_L", 2, 2, seed=0, gain=2.5)
    assert m.load_state_dict(sd, strict=False).missing_keys == []
    o = m.encode_gallery(synth.make_images(N).to(dev), want_f32=False, want_bf16=True, want_raws_f32=False,
                         want_raws_bf16=True)
    feats, raws = o["feats_bf16"], o["raws_bf16"]
    names = [f"img{i:03d}" for i in range(N)]
    ids, mask = synth.make_token_ids(Q)
    ref = torch.randint(0, N, (Q,), generator=torch.Generator().manual_seed(7))
    subset = torch.randint(0, N, (Q, 6), generator=torch.Generator().manual_seed(8))
    whole = RT.GalleryIndex(feats=feats, raws=raws, names=names)
    sc1, ix1, sub1 = RT.query_topk(m, whole, ref, ids, mask, k=k, subset_rows=subset)
    rr1 = RT.rerank_topk(m, whole, ix1.long(), ref, ids, mask, T)

    dist.init_process_group("nccl", device_id=dev)
    lo, hi = RT.shard_range(N, rank, world)
    shard = RT.GalleryIndex(feats=feats[lo:hi].contiguous(), raws=raws[lo:hi].contiguous(), names=names, lo=lo, hi=hi,
                            n_total=N)
    sc, ix, sub = RT.query_topk(m, shard, ref, ids, mask, k=k, subset_rows=subset)
    need = torch.tensor([N - 1, 0, N // 2, 0, 3][: 2 + 3 * (rank % 2)])
    got = RT.fetch_raw_rows(shard, need)
    rr = RT.rerank_topk(m, shard, ix.long(), ref, ids, mask, T)
    torch.cuda.synchronize()
    res = dict(rows_equal=bool(torch.equal(ix, ix1)), scores_equal=bool(torch.equal(sc, sc1)),
               subset_equal=bool(torch.equal(sub, sub1)), fetch_equal=bool(torch.equal(got, raws[need.to(dev)])),
               rerank_equal=bool(torch.equal(rr, rr1)), rerank_reorders=bool(not torch.equal(rr1, ix1.long().cpu())),
               max_dscore=float((sc - sc1).abs().max()))
    torch.save(res, os.path.join(out_dir, f"res.{rank}"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
