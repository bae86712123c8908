"""diffusers 0.21.4 Attention (AttnProcessor2_0 path) / FeedForward / GEGLU / AdaLayerNorm, restated."""
import torch
import torch.nn.functional as F
from torch import nn


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, upcast_softmax=False, added_kv_proj_dim=None, norm_num_groups=None,
                 out_bias=True, scale_qk=True, **unused):
        super().__init__()
        inner_dim = dim_head * heads
        cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention = upcast_attention
        self.scale = dim_head ** -0.5 if scale_qk else 1.0
        self.heads = heads
        self.sliceable_head_dim = heads
        self.added_kv_proj_dim = added_kv_proj_dim
        self.group_norm = None
        self.to_q = nn.Linear(query_dim, inner_dim, bias=bias)
        self.to_k = nn.Linear(cross_attention_dim, inner_dim, bias=bias)
        self.to_v = nn.Linear(cross_attention_dim, inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])

    def set_use_memory_efficient_attention_xformers(self, *a, **k):
        pass

    def head_to_batch_dim(self, tensor, out_dim=3):
        b, s, dim = tensor.shape
        h = self.heads
        tensor = tensor.reshape(b, s, h, dim // h).permute(0, 2, 1, 3)
        if out_dim == 3:
            tensor = tensor.reshape(b * h, s, dim // h)
        return tensor

    def batch_to_head_dim(self, tensor):
        bh, s, d = tensor.shape
        h = self.heads
        return tensor.reshape(bh // h, h, s, d).permute(0, 2, 1, 3).reshape(bh // h, s, d * h)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        b = hidden_states.shape[0]
        query = self.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        key = self.to_k(encoder_hidden_states)
        value = self.to_v(encoder_hidden_states)
        d = key.shape[-1] // self.heads
        query = query.view(b, -1, self.heads, d).transpose(1, 2)
        key = key.view(b, -1, self.heads, d).transpose(1, 2)
        value = value.view(b, -1, self.heads, d).transpose(1, 2)
        hs = F.scaled_dot_product_attention(query, key, value, attn_mask=attention_mask, dropout_p=0.0,
                                            is_causal=False)
        hs = hs.transpose(1, 2).reshape(b, -1, self.heads * d).to(query.dtype)
        hs = self.to_out[0](hs)
        return self.to_out[1](hs)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False):
        super().__init__()
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        assert activation_fn == "geglu"
        self.net = nn.ModuleList([GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out)])

    def forward(self, hidden_states):
        for m in self.net:
            hidden_states = m(hidden_states)
        return hidden_states


class AdaLayerNorm(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("AdaLayerNorm is unused by the InsV2V configs")
