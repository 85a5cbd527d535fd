"""BERTEmbedder — drop-in mirror of frido/modules/encoders/modules.py:85-114 (the stage right before
the hot path): a 32-layer x-transformer *encoder* (frido/modules/x_transformer.py: TransformerWrapper :548, Encoder
:541, Attention :215 with 8 heads x 64, FeedForward :194 with exact GELU, pre-LayerNorm, final LayerNorm) over layout /
caption tokens.  It runs once per batch; its output is the constant `context` of all T x S UNet evaluations.

Same constructor keywords and state-dict keys (`transformer.token_emb.weight`, `transformer.pos_emb.emb.weight`,
`transformer.attn_layers.layers.{i}.{0,1}.*`, `transformer.norm.*`, `transformer.to_logits.*`); the forward is a
libfrido_b200 program: embedding gather, LayerNorm, fused q|k|v GEMM, short-sequence multi-head attention kernel,
output projection + residual, GELU feed-forward — GEMMs on the tcgen05 engine.  The HF tokenizer (`use_tokenizer=True`)
is host-side text processing and needs `bert-base-uncased` on disk; token tensors are always accepted.
"""
import torch
from torch import nn

from . import _lib as L
from . import packing as PK
from . import modules as M
from .program import Program

DEFAULT_DIM_HEAD = 64  # x_transformer.py:19


class _Attention(M._NoForward):
    def __init__(self, dim, heads=8, dim_head=DEFAULT_DIM_HEAD):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Linear(inner, dim)


class _FeedForward(M._NoForward):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = M.Seq(M.Seq(nn.Linear(dim, dim * mult), M.Marker()), M.Marker(), nn.Linear(dim * mult, dim))


class _AbsPos(M._NoForward):
    def __init__(self, dim, max_seq_len):
        super().__init__()
        self.emb = nn.Embedding(max_seq_len, dim)
        nn.init.normal_(self.emb.weight, std=0.02)


class _Encoder(M._NoForward):
    def __init__(self, dim, depth, heads=8):
        super().__init__()
        self.dim, self.depth = dim, depth
        layers = []
        for _ in range(depth):  # layer_types = ('a', 'f') * depth, each [norm, block, residual]
            layers.append(nn.ModuleList([nn.LayerNorm(dim), _Attention(dim, heads), M.Marker()]))
            layers.append(nn.ModuleList([nn.LayerNorm(dim), _FeedForward(dim), M.Marker()]))
        self.layers = nn.ModuleList(layers)


class _TransformerWrapper(M._NoForward):
    def __init__(self, num_tokens, max_seq_len, dim, depth):
        super().__init__()
        self.max_seq_len, self.num_tokens = max_seq_len, num_tokens
        self.token_emb = nn.Embedding(num_tokens, dim)
        nn.init.normal_(self.token_emb.weight, std=0.02)
        self.pos_emb = _AbsPos(dim, max_seq_len)
        self.attn_layers = _Encoder(dim, depth)
        self.norm = nn.LayerNorm(dim)
        self.to_logits = nn.Linear(dim, num_tokens)  # unused for embeddings; kept for checkpoint compatibility


class BERTEmbedder(nn.Module):
    def __init__(self, n_embed, n_layer, vocab_size=30522, max_seq_len=77, device="cuda", use_tokenizer=True,
                 embedding_dropout=0.0, cond_key=""):
        super().__init__()
        self.use_tknz_fn = use_tokenizer
        self.cond_key = cond_key
        self.device = device
        self.max_seq_len = max_seq_len
        self.tknz_fn = None
        self.transformer = _TransformerWrapper(vocab_size, max_seq_len, n_embed, n_layer)
        self._plans = {}

    def invalidate(self):
        self._plans.clear()

    def _tokenize(self, text):
        if self.tknz_fn is None:
            from transformers import BertTokenizerFast  # needs bert-base-uncased on disk (no network here)
            self.tknz_fn = BertTokenizerFast.from_pretrained("bert-base-uncased")
        enc = self.tknz_fn(text, truncation=True, max_length=self.max_seq_len, return_length=True,
                           return_overflowing_tokens=False, padding="max_length", return_tensors="pt")
        return enc["input_ids"]

    @torch.no_grad()
    def forward(self, text, return_token=False):
        if torch.is_tensor(text):
            tokens = text.long()
        elif self.use_tknz_fn:
            tokens = self._tokenize(text)
        else:
            tokens = (text[self.cond_key] if self.cond_key != "" else text).long()
        dev = self.transformer.norm.weight.device
        if dev.type != "cuda":
            raise L.FridoError("BERTEmbedder runs on a CUDA device only (no CPU path)")
        tokens = tokens.to(dev)
        B, Lseq = tokens.shape
        assert Lseq <= self.max_seq_len
        plan = self._plans.get((B, Lseq))
        if plan is None:
            plan = _EmbedPlan(self, B, Lseq, dev)
            self._plans[(B, Lseq)] = plan
        plan.repack()
        plan.tokens.copy_(tokens)
        plan.prog.run()
        z = plan.out.clone()
        return (z, tokens) if return_token else z

    def encode(self, text):
        return self(text)


class _EmbedPlan:
    def __init__(self, enc: BERTEmbedder, B, Lseq, dev):
        tw = enc.transformer
        D = tw.token_emb.weight.shape[1]
        self.packers = []
        P = self.prog = Program(dev, "bert_embedder")
        self.tokens = torch.zeros(B, Lseq, dtype=torch.int64, device=dev)
        M_ = B * Lseq
        x = P.buf(B, Lseq, D)
        P.embed_tokens(self.tokens, self._p(lambda: PK.copy(tw.token_emb.weight)),
                       self._p(lambda: PK.copy(tw.pos_emb.emb.weight)), x, B=B, Lseq=Lseq, D=D)
        for norm, blk, _ in tw.attn_layers.layers:
            ln = P.buf(M_, D)
            P.layernorm(x, self._v(norm.weight), self._v(norm.bias), ln, rows=M_, Cdim=D, round_tf32=P.R)
            if isinstance(blk, _Attention):
                H, Dh = blk.heads, blk.dim_head
                inner = H * Dh
                wqkv = self._p(lambda b=blk: PK.cat_rows([b.to_q.weight, b.to_k.weight, b.to_v.weight]))
                qkv = P.buf(M_, 3 * inner)
                P.linear(ln, wqkv, qkv, M=M_, K=D, N=3 * inner, tag="enc.qkv")
                att = P.buf(M_, inner)
                P.mha_small(qkv, att, B=B, Lseq=Lseq, H=H, Dh=Dh, scale=float(Dh) ** -0.5)
                y = P.buf(B, Lseq, D)
                P.linear(att, self._v(blk.to_out.weight), y, M=M_, K=inner, N=D, bias=self._v(blk.to_out.bias), res=x,
                         round_tf32=0, tag="enc.attn_out")
                P.release(qkv); P.release(att)
            else:
                lin0, lin2 = blk.net[0][0], blk.net[2]
                hdim = lin0.weight.shape[0]
                hbuf = P.buf(M_, hdim)
                P.linear(ln, self._v(lin0.weight), hbuf, M=M_, K=D, N=hdim, bias=self._v(lin0.bias), act=L.ACT_GELU,
                         round_tf32=P.R, tag="enc.ff0")
                y = P.buf(B, Lseq, D)
                P.linear(hbuf, self._v(lin2.weight), y, M=M_, K=hdim, N=D, bias=self._v(lin2.bias), res=x, tag="enc.ff2")
                P.release(hbuf)
            P.release(ln); P.release(x)
            x = y
        self.out = torch.zeros(B, Lseq, D, dtype=torch.float32, device=dev)
        P.layernorm(x, self._v(tw.norm.weight), self._v(tw.norm.bias), self.out, rows=M_, Cdim=D)
        self.repack()

    def _p(self, fn):
        dst = fn().contiguous()
        self.packers.append((dst, fn))
        return dst

    def _v(self, prm):
        return self._p(lambda: PK.copy(prm))

    def repack(self):
        """Weights are re-read on every call (the encoder may be trainable / EMA-swapped)."""
        for dst, fn in self.packers:
            PK.place(fn().view(1, -1), dst.view(1, -1))
        self.prog.prepare_weights()
