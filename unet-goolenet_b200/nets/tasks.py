"""Parameter shells of the reference's transformer bottleneck (分割/nets/tasks.py:46-231).

Only the module/parameter tree is defined here (names and shapes as the reference's, so checkpoints load
strictly); the arithmetic is lowered to engine ops by ugnet_b200.lower.  Calling these sub-modules directly is
not supported — the enclosing UNetTaskAligWeight.forward runs the whole network on the engine."""
import torch
import torch.nn as nn


def _engine_only(name):
    def forward(self, *a, **k):
        raise NotImplementedError(f"{name} runs only as part of UNetTaskAligWeight.forward on the ugnet engine")
    return forward


class FeedForward(nn.Module):  # tasks.py:46-57
    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))
    forward = _engine_only("FeedForward")


class Cross_Attention(nn.Module):  # tasks.py:58-97
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0, softmax=True):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.scale, self.softmax = heads, dim ** -0.5, softmax
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))
    forward = _engine_only("Cross_Attention")


class Attention(nn.Module):  # tasks.py:121-148
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.scale = heads, dim ** -0.5
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))
    forward = _engine_only("Attention")


class Conv2dReLU(nn.Sequential):  # tasks.py:98-120
    def __init__(self, in_channels, out_channels, kernel_size, padding=0, stride=1, use_batchnorm=True):
        super().__init__(nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                   bias=not use_batchnorm),
                         nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))


class Multi_Attention(nn.Module):  # tasks.py:149-184
    def __init__(self, dim, heads, dim_head, mlp_dim, dropout, softmax=True):
        super().__init__()
        self.attention1 = Attention(dim, heads=heads, dim_head=dim_head, dropout=0)
        self.attention2 = Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)
        self.cross_attention_cl = Cross_Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout,
                                                  softmax=softmax)
        self.cross_attention_seg = Cross_Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout,
                                                   softmax=softmax)
        self.x_att_norm = nn.LayerNorm(dim)
        self.m_att_norm = nn.LayerNorm(dim)
        self.x_mlp_norm = nn.LayerNorm(dim)
        self.m_mlp_norm = nn.LayerNorm(dim)
        self.x_feed = FeedForward(dim, mlp_dim, dropout=dropout)
        self.m_feed = FeedForward(dim, mlp_dim, dropout=dropout)
    forward = _engine_only("Multi_Attention")


class TransformerDecoder(nn.Module):  # tasks.py:188-231
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout, decoder_pos_size, softmax=True):
        super().__init__()
        if depth != 1:
            raise NotImplementedError("the engine lowers depth == 1, the only depth the reference instantiates "
                                      "(basicUnet.py:397-399)")
        self.conv_cl = Conv2dReLU(dim, dim, kernel_size=3, padding=1, use_batchnorm=True)
        self.conv_seg = Conv2dReLU(dim, dim, kernel_size=3, padding=1, use_batchnorm=True)
        self.layers = nn.ModuleList([Multi_Attention(dim, heads=heads, dim_head=dim_head, mlp_dim=mlp_dim,
                                                     dropout=dropout, softmax=softmax) for _ in range(depth)])
        shape = (1, dim, decoder_pos_size, decoder_pos_size)
        self.pos_embedding_decoder_cl = nn.Parameter(torch.zeros(shape))
        self.pos_embedding_decoder_seg = nn.Parameter(torch.zeros(shape))
    forward = _engine_only("TransformerDecoder")
