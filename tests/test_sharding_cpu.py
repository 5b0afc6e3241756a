"""Multi-rank host logic on CPU: frame sharding + one all_gather + one all_reduce (gloo,
world_size 2), and the per-frame .pth output format."""
import os
import socket
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from findnpropagate_b200 import extract


def _fake_compute(frames):
    """Deterministic stand-in for engine.run on CPU: proposals derived from the frame id."""
    out, rec = [], {k: 0 for k in extract.recall_keys()}
    for f in frames:
        k = f % 4
        rng = np.random.default_rng(f)
        out.append(dict(pred_boxes=rng.normal(size=(k, 7)).astype(np.float32),
                        pred_scores=rng.uniform(size=k).astype(np.float32),
                        pred_labels=rng.integers(1, 11, k).astype(np.int32)))
        rec["gt"] += 3
        rec["rcnn_0.3"] += k % 3
        rec["rcnn_7unknown_0.5"] += 1
    return dict(frames=out, recall=rec)


def _worker(rank, world, port, n_frames, folder, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = list(range(n_frames))
    merged, total, ar = extract.extract(frames, _fake_compute, folder=folder, batch_frames=3, rank=rank, world=world,
                                        frame_ids=["n%03d.pcd.bin" % i for i in frames])
    q.put((rank, [m["pred_boxes"].tobytes() for m in merged], total, ar))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_two_rank_extract_matches_single_rank():
    n = 11
    single, total1, ar1 = extract.extract(list(range(n)), _fake_compute, batch_frames=4)
    with tempfile.TemporaryDirectory() as d:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n, d, q)) for r in range(2)]
        [p.start() for p in procs]
        res = [q.get(timeout=120) for _ in range(2)]
        [p.join(60) for p in procs]
        assert all(p.exitcode == 0 for p in procs)
        for rank, boxes, total, ar in res:
            assert boxes == [m["pred_boxes"].tobytes() for m in single]       # dataset order restored
            assert total == total1 and ar == ar1
        files = sorted(os.listdir(d))
        assert len(files) == n and files[0] == "n000_pcd_bin.pth"
        out = torch.load(os.path.join(d, files[5]), map_location="cpu")
        assert isinstance(out, list) and len(out) == 1
        assert out[0]["pred_boxes"].dtype == torch.float32 and out[0]["pred_boxes"].shape[1] == 7
        assert out[0]["pred_labels"].dtype == torch.int32 and out[0]["pred_scores"].dtype == torch.float32
        assert out[0]["pred_boxes"].shape[0] == 5 % 4


def test_shard_rule_and_pack_roundtrip():
    assert extract.shard_indices(10, 1, 4) == [1, 5, 9]
    preds = _fake_compute([1, 2, 3, 7])["frames"]
    pack, cnt = extract.pack_proposals(preds, 5)
    back = extract.unpack_proposals(pack, cnt)
    for a, b in zip(preds, back):
        assert np.array_equal(a["pred_boxes"], b["pred_boxes"]) and np.array_equal(a["pred_labels"], b["pred_labels"])


def test_glip_feeder_roundtrip(tmp_path):
    """PreprocessedGLIP reads the pickled BoxList file + COCO meta json of the reference."""
    import json
    from findnpropagate_b200 import proposer
    proposer.install_boxlist_shim()
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    lists, images = [], []
    for i in range(12):
        bl = BoxList(torch.rand(3, 4) * 100, (1600, 900))
        bl.extra_fields = dict(scores=torch.rand(3), labels=torch.randint(1, 11, (3,)))
        lists.append(bl)
        images.append(dict(token="tok%d" % (i // 6), file_name="cam%d_%d.jpg" % (i % 6, i // 6)))
    torch.save(lists, tmp_path / "pred.pth")
    json.dump(dict(images=images, categories=[]), open(tmp_path / "meta.json", "w"))
    feeder = proposer.PreprocessedGLIP(str(tmp_path / "pred.pth"), str(tmp_path / "meta.json"))
    bd = dict(batch_size=2, image_paths=[["cam%d_%d.jpg" % (c, b) for c in range(6)] for b in range(2)],
              metadata=[dict(token="tok0"), dict(token="tok1")])
    boxes, labels, scores, bidx, cam = feeder(bd)
    assert boxes.shape == (36, 4) and labels.shape == (36,) and cam.tolist()[:6] == [0, 0, 0, 1, 1, 1]
    assert bidx.tolist() == [0] * 18 + [1] * 18


def test_pseudo_file_reader_roundtrip(tmp_path):
    """extract.save_frame -> pseudo_loader.load_pseudos: the reference's file format and the (K,8)
    layout of PseudoLoader.load_pseudos (pseudo_loader.py:561-679); missing file -> empty."""
    from findnpropagate_b200 import extract, pseudo_loader
    rng = np.random.default_rng(0)
    pred = dict(pred_boxes=rng.normal(size=(5, 7)).astype(np.float32), pred_scores=rng.random(5).astype(np.float32),
                pred_labels=np.array([1, 9, 3, 9, 10], np.int32))
    path = extract.save_frame(str(tmp_path), "n015-2018.pcd.bin", pred)
    assert path.endswith("n015-2018_pcd_bin.pth")
    boxes, scores = pseudo_loader.load_pseudos(tmp_path, "n015-2018.pcd.bin")
    assert boxes.shape == (5, 8) and boxes.dtype == np.float32
    assert np.array_equal(boxes[:, :7], pred["pred_boxes"]) and np.array_equal(boxes[:, 7], pred["pred_labels"])
    assert np.array_equal(scores, pred["pred_scores"])
    b9, s9 = pseudo_loader.load_pseudos(tmp_path, "n015-2018.pcd.bin", labels={9})
    assert b9.shape == (2, 8) and np.all(b9[:, 7] == 9)
    e, es = pseudo_loader.load_pseudos(tmp_path, "missing")
    assert e.shape == (0, 8) and es.shape == (0,)
