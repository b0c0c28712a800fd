"""Checkpoint compatibility: the modules carry the state-dict keys the reference produces after
``to_hetero_old`` (``lightning_model.py:371-421``), and a Lightning-format ``.ckpt`` (``state_dict`` + ``hyper_parameters``)
loads through ``load_from_checkpoint`` (``main.py:216-233,319-334``).  CPU only: no kernel is launched."""
import argparse
import re

import torch

TARGET_RELS = ["count__union_triangle__count", "count__union_tride__count", "count__union_triangle__canonical",
               "count__union_tride__canonical", "canonical__union_triangle__count", "canonical__union_tride__count"]
QUERY_RELS = ["union_node__union_triangle__union_node", "union_node__union_tride__union_node"]


def _expected_neighborhood_keys(layers=8):
    keys = []
    for prefix, types, rels, anchor in (("emb_model", ["count", "canonical"], TARGET_RELS, True),
                                        ("emb_model_query", ["union_node"], QUERY_RELS, True)):
        for t in types:
            keys += [f"{prefix}.gnn_core.pre_mp.0.{t}.{p}" for p in ("weight", "bias")]
        for l in range(layers):
            for r in rels:
                keys += [f"{prefix}.gnn_core.convs.{l}.{r}.lin.{p}" for p in ("weight", "bias")]
            for t in types:
                keys += [f"{prefix}.gnn_core.updates.{l}.{t}.{p}" for p in ("weight", "bias")]
        keys += [f"{prefix}.anchor_mlp.0.{p}" for p in ("weight", "bias")]
        for i in (0, 3, 5, 7):
            keys += [f"{prefix}.post_mp.{i}.{p}" for p in ("weight", "bias")]
    keys += [f"count_model.{i}.{p}" for i in (0, 2) for p in ("weight", "bias")]
    return keys


def test_neighborhood_state_dict_keys_follow_to_hetero_old():
    from desco_b200.lightning_model import NeighborhoodCountingModel

    got = set(NeighborhoodCountingModel().state_dict().keys())
    assert got == set(_expected_neighborhood_keys())


def test_gossip_state_dict_keys():
    from desco_b200.lightning_model import GossipCountingModel

    got = set(GossipCountingModel().state_dict().keys())
    exp = {f"emb_model.gnn_core.pre_mp.0.{p}" for p in ("weight", "bias")}
    for l in (0, 1):
        for m in ("lin_com", "lin_update", "lin_gate.0", "lin_gate.2"):
            exp |= {f"emb_model.gnn_core.convs.{l}.{m}.{p}" for p in ("weight", "bias")}
    exp |= {f"emb_model.anchor_mlp.0.{p}" for p in ("weight", "bias")}
    exp |= {f"emb_model.post_mp.{i}.{p}" for i in (0, 3, 5, 7) for p in ("weight", "bias")}
    assert got == exp
    assert all(re.fullmatch(r"[a-z_.0-9]+", k) for k in got)


def test_lightning_checkpoint_round_trip(tmp_path):
    from desco_b200.lightning_model import GossipCountingModel, NeighborhoodCountingModel

    torch.manual_seed(3)
    src = NeighborhoodCountingModel()
    args = argparse.Namespace(conv_type="SAGE", layer_num=8, hidden_dim=64, input_dim=1, dropout=0.0, use_hetero=True,
                              use_tconv=True, depth=4, lr=1e-4, weight_decay=0.0, batch_size=512)  # no use_canonical: old ckpt
    path = tmp_path / "neigh.ckpt"
    torch.save({"state_dict": src.state_dict(), "hyper_parameters": {"input_dim": 1, "hidden_dim": 64, "args": args},
                "epoch": 7, "pytorch-lightning_version": "1.6.4"}, path)
    dst = NeighborhoodCountingModel.load_from_checkpoint(str(path))
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    assert dst.args.layer_num == 8

    bad = argparse.Namespace(**{**vars(args), "use_tconv": False})
    torch.save({"state_dict": src.state_dict(), "hyper_parameters": {"input_dim": 1, "hidden_dim": 64, "args": bad}}, path)
    try:
        NeighborhoodCountingModel.load_from_checkpoint(str(path))
        raise AssertionError("a non-SHMP checkpoint must be refused")
    except NotImplementedError:
        pass

    gsrc = GossipCountingModel()
    gargs = argparse.Namespace(conv_type="GOSSIP", layer_num=2, hidden_dim=64, dropout=0.01, use_hetero=False, lr=1e-3,
                               weight_decay=0.0, batch_size=256)
    gpath = tmp_path / "gossip.ckpt"
    torch.save({"state_dict": gsrc.state_dict(),
                "hyper_parameters": {"input_dim": 1, "hidden_dim": 64, "args": gargs, "emb_channels": 64, "input_pattern_emb": True}},
               gpath)
    gdst = GossipCountingModel.load_from_checkpoint(str(gpath))
    for (k, a), (_, b) in zip(gsrc.state_dict().items(), gdst.state_dict().items()):
        assert torch.equal(a, b), k
