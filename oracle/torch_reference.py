"""TEST / BASELINE INFRASTRUCTURE -- not part of the product path (only tests/, bench.py's baseline legs and
__graft_entry__.smoke() may import oracle/).

Stock-PyTorch restatement of the reference module on ``nn.Transformer`` (model.py:60-170): the same library calls the
reference makes (nn.Linear, nn.InstanceNorm1d on a [S,N,E] tensor, nn.Transformer(activation="gelu", dropout=0), F.sigmoid),
with the reference's parameter names so a state_dict of the drop-in module (or a reference ``.pth``) loads unchanged.  It is
what bench.py times as

  * ``gpu_baseline``: "stock PyTorch on the B200 running the same module" (SURVEY.md section 2 / 8(d)), fp32 and bf16 autocast;
  * ``cpu_baseline`` / ``--impl reference``: the reference's CPU path on the box's host cores (/root/reference does not travel
    to the GPU box, and model.py hard-codes 54 keypoints in its last ``view`` -- model.py:163 -- so the K = 71 benchmark shape
    needs this restatement's generalised reshape anyway).

Pinned against outputs of the reference itself: tests/test_oracle_golden.py::test_torch_reference_matches_reference_goldens.
"""
import math

import torch
import torch.nn as nn


class _Gate(nn.Module):
    """model.py:11-22 -- fc3(fc1(x) * sigmoid(fc2(x)))."""

    def __init__(self, dim):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)

    def forward(self, x):
        return self.fc3(self.fc1(x) * torch.sigmoid(self.fc2(x)))


class _TrigTable(nn.Module):
    """model.py:24-50 -- x + pe[:S] with the sin / cos table as a buffer named ``pos_encoding``."""

    def __init__(self, dim, max_len):
        super().__init__()
        pos = torch.arange(max_len, dtype=torch.float).unsqueeze(1)
        freq = torch.exp(torch.arange(0, dim, 2).float() * (-math.log(10000.0)) / dim)
        table = torch.zeros(max_len, dim)
        table[:, 0::2], table[:, 1::2] = torch.sin(pos * freq), torch.cos(pos * freq)
        self.register_buffer("pos_encoding", table.unsqueeze(1))

    def forward(self, x):
        return x + self.pos_encoding[:x.size(0)]


class StockCompleter(nn.Module):
    def __init__(self, input_size, hidden_dim, num_layers, num_heads, max_len=2048):
        super().__init__()
        self.input_size, self.num_heads = input_size, num_heads
        self.input_embedding = nn.Linear(input_size, hidden_dim)
        self.filled_embedding = nn.Linear(input_size, hidden_dim)
        self.input_norm1 = nn.InstanceNorm1d(hidden_dim)      # applied to [S,N,E]: normalises over E per token
        self.filled_norm1 = nn.InstanceNorm1d(hidden_dim)
        self.trig_input_positional_encoder = _TrigTable(hidden_dim, max_len)
        self.trig_filled_positional_encoder = _TrigTable(hidden_dim, max_len)
        self.learned_input_positional_encoder = nn.Parameter(torch.rand(1, 1, hidden_dim))
        self.learned_filled_positional_encoder = nn.Parameter(torch.rand(1, 1, hidden_dim))
        self.swiGlu_input_prev = _Gate(hidden_dim)
        self.swiGlu_filled_prev = _Gate(hidden_dim)
        self.transformer = nn.Transformer(d_model=hidden_dim, nhead=num_heads, num_encoder_layers=num_layers,
                                          num_decoder_layers=num_layers, activation="gelu", dropout=0.0)
        self.swiGlu_decoded = _Gate(hidden_dim)
        self.norm2 = nn.InstanceNorm1d(hidden_dim)
        self.fc_final = nn.Linear(hidden_dim, input_size)

    def _branch(self, frames, embed, norm, trig, learned, gate):
        seq_first = frames.flatten(start_dim=2).float().transpose(0, 1)     # [S,N,2K]
        emb = embed(seq_first)
        return emb, gate(trig(norm(emb)) + learned)

    def forward(self, inputs, filled, src_pad_mask=None, src_mask=None, tgt_mask=None):
        """inputs / filled [N,S,K,2]; src_pad_mask [N,S] float (ADDED to the logits); src_mask / tgt_mask additive,
        [S,S] or [N*heads,S,S].  Returns [N,S,K,2]."""
        _, x = self._branch(inputs, self.input_embedding, self.input_norm1, self.trig_input_positional_encoder,
                            self.learned_input_positional_encoder, self.swiGlu_input_prev)
        filled_emb, y = self._branch(filled, self.filled_embedding, self.filled_norm1, self.trig_filled_positional_encoder,
                                     self.learned_filled_positional_encoder, self.swiGlu_filled_prev)
        dec = self.transformer(x, y, src_mask=src_mask, tgt_mask=tgt_mask, src_key_padding_mask=src_pad_mask)
        dec = self.norm2(self.swiGlu_decoded(dec) + filled_emb)
        dec = dec * torch.sigmoid(dec)
        out = self.fc_final(dec.transpose(0, 1))
        return out.reshape(out.shape[0], out.shape[1], self.input_size // 2, 2)


def repeat_inc_masks(frame_mask, heads):
    """Batched model.get_mask(mask, T, "repeat-inc") (model.py:193-202): [N,S] 0/1 -> [N*heads,S,S] float, -inf iff j > i and
    mask[n,j] == 1.  Vectorised here (the reference's Python double loop costs 17.6 ms per sequence at T = 64)."""
    n, s = frame_mask.shape
    i = torch.arange(s, device=frame_mask.device).view(1, s, 1)
    j = torch.arange(s, device=frame_mask.device).view(1, 1, s)
    m = torch.zeros(n, s, s, device=frame_mask.device).masked_fill((j > i) & (frame_mask.view(n, 1, s) == 1), float("-inf"))
    return m.repeat_interleave(heads, dim=0)


def get_mask_loop(mask, size):
    """model.get_mask(mask, size, "repeat-inc") as the reference computes it for ONE sequence: the Python double loop
    (model.py:193-202).  Used by the A1-faithful batch-1 CPU baseline, whose cost it dominates."""
    m = mask.clone().reshape(1, size).repeat(size, 1)
    m = torch.where(m == 1, torch.tensor(float("-inf")), m)
    for i in range(size):
        for j in range(i + 1):
            m[i, j] = 0.0
    return m


def train_step(model, opt, inputs, sota, mask, autocast_dtype=None):
    """A1_train.py:91-135 for a batch: slices, masks, forward, MSELoss, zero_grad / backward / Adam."""
    x, x_dec = inputs[:, :-1], inputs[:, 1:]
    x_mask, y_mask = mask[:, :-1], mask[:, 1:]
    src_mask = repeat_inc_masks(x_mask, model.num_heads)
    tgt_mask = repeat_inc_masks(y_mask, model.num_heads)
    if autocast_dtype is not None:
        with torch.autocast(device_type=inputs.device.type, dtype=autocast_dtype):
            pred = model(x, x_dec, src_pad_mask=x_mask, src_mask=src_mask, tgt_mask=tgt_mask)
    else:
        pred = model(x, x_dec, src_pad_mask=x_mask, src_mask=src_mask, tgt_mask=tgt_mask)
    loss = torch.nn.functional.mse_loss(pred.float(), sota)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss
