"""ORACLE / test infrastructure: golden vectors for the text encoder, produced by EXECUTING the unmodified
reference `src/models/text_encoder.py:TextEncoder` (build container only).

    python -m oracle.make_golden_text            # seconds

The reference's constructor downloads `BertTokenizer` / `BertModel` with `from_pretrained`; there is no network here, so those two
third-party loaders are replaced for the duration of the call by stand-ins that return (a) a BERT-mini-shaped `transformers.BertModel`
built from a `BertConfig` under a fixed seed (random weights, the real architecture and forward code of the installed transformers) and
(b) a tokenizer stub that returns pre-drawn token ids with a right-padding attention mask.  Everything downstream -- the reference's own
`__init__` (fine-tuning flags, projection, LayerNorm) and `forward` (bert -> projection -> layer_norm) -- runs unmodified.

Output: tests/golden/text_encoder.pt -- inputs (ids, mask), outputs for hidden_dim 256 (no projection, the stage-2 configuration) and
hidden_dim 384 (with projection), and the requires_grad pattern per fine-tuning strategy.  The tests rebuild the same weights from the
same seed (11 M parameters are not shipped).
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

from . import ref_loader

OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"
BERT_MINI = dict(hidden_size=256, num_hidden_layers=4, num_attention_heads=4, intermediate_size=1024)


def build_bert(seed: int):
    """BERT-mini-shaped BertModel with seeded random weights (shared with tests/test_text_encoder.py)."""
    from transformers import BertConfig, BertModel
    torch.manual_seed(seed)
    return BertModel(BertConfig(**BERT_MINI)).eval()


def text_inputs(batch: int, length: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, 30000, (batch, length), generator=g)
    lens = torch.randint(max(1, length // 3), length + 1, (batch,), generator=g)
    lens[0] = length                                         # the longest sample defines the padded length, as padding=True does
    mask = (torch.arange(length)[None, :] < lens[:, None]).long()
    ids = ids * mask                                         # [PAD] = 0 on the masked suffix
    return ids, mask


class _Tok:
    def __init__(self, ids, mask):
        self.ids, self.mask = ids, mask

    def __call__(self, text_list, **kw):
        assert kw.get("padding") is True and kw.get("return_tensors") == "pt"
        return {"input_ids": self.ids, "attention_mask": self.mask, "token_type_ids": torch.zeros_like(self.ids)}


def main():
    torch.set_num_threads(8)
    ref_loader.install_stubs()
    root = ref_loader.REFERENCE_ROOT
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    import src.models.text_encoder as te  # type: ignore
    golden = {"bert_seed": 7, "cases": {}, "requires_grad": {}}
    for name, (hidden, batch, length, seed) in {"h256_b3_l32": (256, 3, 32, 91), "h384_b2_l19": (384, 2, 19, 92)}.items():
        ids, mask = text_inputs(batch, length, seed)
        real_tok, real_bert = te.BertTokenizer.from_pretrained, te.BertModel.from_pretrained
        te.BertTokenizer.from_pretrained = staticmethod(lambda *_a, **_k: _Tok(ids, mask))
        te.BertModel.from_pretrained = staticmethod(lambda *_a, **_k: build_bert(7))
        try:
            torch.manual_seed(13)                            # projection / layer_norm initialisation
            enc = te.TextEncoder(model_name="prajjwal1/bert-mini", hidden_dim=hidden, finetune_strategy="minimal").eval()
            with torch.no_grad():
                out = enc(["x"] * batch)
            for strat in ("none", "minimal", "partial", "full"):
                e2 = te.TextEncoder(model_name="prajjwal1/bert-mini", hidden_dim=hidden, finetune_strategy=strat)
                golden["requires_grad"][f"{hidden}_{strat}"] = [k for k, p in e2.named_parameters() if p.requires_grad]
        finally:
            te.BertTokenizer.from_pretrained, te.BertModel.from_pretrained = real_tok, real_bert
        golden["cases"][name] = {"hidden": hidden, "ids": ids, "mask": mask, "head_seed": 13, "output": out.clone(),
                                 "keys": list(enc.state_dict().keys())}
        print(f"{name}: out {tuple(out.shape)} std {out.std():.4f}", flush=True)
    torch.save(golden, OUT / "text_encoder.pt")
    print("wrote text_encoder.pt")


if __name__ == "__main__":
    main()
