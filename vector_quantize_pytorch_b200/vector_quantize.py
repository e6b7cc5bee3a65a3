"""`VectorQuantize` — drop-in for the reference module (vector_quantize_pytorch.py:802-1403) on the
default path: heads=1, no mask, Euclidean or cosine codebook with EMA updates.

forward(x) -> (quantize [x.dtype, x.shape], embed_ind [int64, x.shape[:-1]], loss [fp32 scalar]).
Layout handling, projections, STE / rotation trick stay PyTorch glue; the search, gather, commitment
loss and EMA run in the sm_100a kernels.  Unsupported constructor / forward options raise.
"""
from __future__ import annotations

from collections import namedtuple

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .codebook import Codebook, _unsupported

LossBreakdown = namedtuple("LossBreakdown", ["commitment", "codebook_diversity", "orthogonal_reg", "inplace_optimize"])


def _is_distributed():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _safe_div(num, den, eps=1e-6):
    return num / den.clamp(min=eps)


def straight_through(src, tgt):  # vqp:282-283
    return src + (tgt - src).detach()


class _RotateTo(torch.autograd.Function):
    """Rotation-trick gradient estimator (arXiv:2410.06424 §4.2; reference vqp:287-318) on the sm_100a kernels: the forward
    value (numerically the quantized vector) and d/d src — the direction / norm factors are constants of the backward pass,
    exactly like the reference's `.detach()`s."""

    @staticmethod
    def forward(ctx, src, tgt):
        ctx.save_for_backward(src, tgt)
        return ops.rotate(src.detach(), tgt.detach())

    @staticmethod
    def backward(ctx, grad_out):
        src, tgt = ctx.saved_tensors
        return ops.rotate(src.detach(), tgt.detach(), grad_out.to(src.dtype)), None


def rotate_to(src, tgt):
    return _RotateTo.apply(src, tgt)


def host_chunk_bounds(N: int, n_chunks: int):
    """Row boundaries [0, b1, ..., N] of the chunks `forward_host` streams through the GPU.

    Chunk sizes ramp up and down (1, 2, 3, 3, ..., 3, 2, 1 units): the first upload and the last download are not
    overlapped with anything, so they are kept short; the middle chunks are large to amortise per-chunk costs.
    Interior boundaries are multiples of 256 rows (whole CTA-pair tiles, 16-byte aligned row offsets)."""
    n_chunks = max(1, int(n_chunks))
    wts = [min(c + 1, n_chunks - c, 3) for c in range(n_chunks)]
    unit = N / sum(wts)
    bounds, acc = [0], 0.0
    for wgt in wts:
        acc += wgt * unit
        b_ = min(N, -(-int(round(acc)) // 256) * 256)
        if b_ > bounds[-1]:
            bounds.append(b_)
    bounds[-1] = N
    return bounds


class VectorQuantize(nn.Module):
    def __init__(
        self,
        dim,
        codebook_size,
        codebook_dim=None,
        heads=1,
        separate_codebook_per_head=False,
        decay=0.8,
        eps=1e-5,
        freeze_codebook=False,
        kmeans_init=False,
        kmeans_iters=10,
        sync_kmeans=True,
        use_cosine_sim=False,
        layernorm_after_project_in=False,
        threshold_ema_dead_code=0,
        channel_last=True,
        accept_image_fmap=False,
        accept_3d_fmap=False,
        commitment_weight=1.,
        commitment_use_cross_entropy_loss=False,
        orthogonal_reg_weight=0.,
        orthogonal_reg_active_codes_only=False,
        orthogonal_reg_max_codes=None,
        codebook_diversity_loss_weight=0.,
        codebook_diversity_temperature=100.,
        stochastic_sample_codes=False,
        sample_codebook_temp=1.,
        straight_through=False,
        rotation_trick=None,
        directional_reparam=False,
        directional_reparam_variance=5e-3,
        sync_codebook=None,
        sync_affine_param=False,
        ema_update=None,
        vq_bridge=None,
        manual_ema_update=False,
        learnable_codebook=None,
        in_place_codebook_optimizer=None,
        manual_in_place_optimizer_update=False,
        affine_param=False,
        affine_param_batch_decay=0.99,
        affine_param_codebook_decay=0.9,
        sync_update_v=0.,
        return_zeros_for_masked_padding=True,
        route_gradients_to_input=True,
    ):
        super().__init__()
        # ---- options outside the accelerated path fail loudly

        if directional_reparam or vq_bridge is not None or learnable_codebook:
            _unsupported("directional_reparam / vq_bridge / learnable_codebook")
        if in_place_codebook_optimizer is not None:
            _unsupported("in_place_codebook_optimizer")
        if affine_param:
            _unsupported("affine_param")
        if stochastic_sample_codes or straight_through:
            _unsupported("stochastic_sample_codes / gumbel straight_through")
        if commitment_use_cross_entropy_loss or orthogonal_reg_weight > 0 or codebook_diversity_loss_weight > 0:
            _unsupported("cross-entropy / orthogonal / diversity losses (they need the N x K distance matrix)")
        if sync_update_v > 0:
            _unsupported("sync_update_v")
        ema_update = True if ema_update is None else ema_update  # vqp:854
        if not ema_update:
            # a frozen, non-learnable codebook is still a valid use of the search kernels
            pass
        rotation_trick = (dim > 1) if rotation_trick is None else rotation_trick  # vqp:856

        self.dim = dim
        self.heads = heads
        self.separate_codebook_per_head = separate_codebook_per_head
        codebook_dim = dim if codebook_dim is None else codebook_dim
        codebook_input_dim = codebook_dim * heads
        requires_projection = codebook_input_dim != dim
        if requires_projection:  # vqp:867-874
            layers = [nn.Linear(dim, codebook_input_dim)]
            if layernorm_after_project_in:
                layers.append(nn.LayerNorm(codebook_input_dim))
            self.project_in = layers[0] if len(layers) == 1 else nn.Sequential(*layers)
            self.project_out = nn.Linear(codebook_input_dim, dim)
        else:
            self.project_in = nn.Identity()
            self.project_out = nn.Identity()
        self.has_projections = requires_projection

        self.eps = eps
        self.has_commitment_loss = commitment_weight > 0.
        self.commitment_weight = commitment_weight
        self.learnable_codebook = False
        self.rotation_trick = rotation_trick
        self.route_gradients_to_input = route_gradients_to_input

        if sync_codebook is None:  # vqp:925-926
            sync_codebook = _is_distributed()

        self.use_cosine_sim = use_cosine_sim
        self._codebook = Codebook(
            dim=codebook_dim,
            num_codebooks=heads if separate_codebook_per_head else 1,  # vqp:931
            codebook_size=codebook_size,
            decay=decay,
            eps=eps,
            threshold_ema_dead_code=threshold_ema_dead_code,
            kmeans_init=kmeans_init,
            kmeans_iters=kmeans_iters,
            use_ddp=sync_codebook,
            sync_kmeans=sync_kmeans,
            sample_codebook_temp=sample_codebook_temp,
            ema_update=ema_update,
            manual_ema_update=manual_ema_update,
            use_cosine_sim=use_cosine_sim,
        )
        self.codebook_size = codebook_size
        self.accept_image_fmap = accept_image_fmap
        self.accept_3d_fmap = accept_3d_fmap
        self.channel_last = channel_last
        self.register_buffer("zero", torch.tensor(0.), persistent=False)  # vqp:970
        self.return_zeros_for_masked_padding = return_zeros_for_masked_padding
        self.freeze_codebook = freeze_codebook

    # ------------------------------------------------------------------ reference surface
    @property
    def ema_update(self):
        return self._codebook.ema_update

    @property
    def codebook(self):  # vqp:982-989
        if self.separate_codebook_per_head:
            return self._codebook.embed
        return self._codebook.embed[0]

    @codebook.setter
    def codebook(self, codes):  # vqp:991-996
        self._codebook.embed.copy_(codes if self.separate_codebook_per_head else codes.unsqueeze(0))

    def get_codes_from_indices(self, indices):  # vqp:998-1018
        if self.separate_codebook_per_head:   # 'b * h' indices -> every head gathers from its own codebook -> 'b * (h d)'
            codes = torch.cat([ops.decode(self._codebook.embed[h], indices[..., h:h + 1].contiguous()) for h in range(self.heads)], dim=-1)
        else:
            codes = ops.decode(self.codebook, indices.unsqueeze(-1))
        if not self.channel_last or self.accept_image_fmap or self.accept_3d_fmap:
            codes = codes.movedim(-1, 1)
        return codes

    def get_output_from_indices(self, indices):  # vqp:1020-1022
        return self.project_out(self.get_codes_from_indices(indices))

    def expire_codes_(self, x):
        self._codebook.expire_codes_(self._codebook.transform_input(x))

    def update_indices(self, x, indices, mask=None):  # vqp:1056-1091
        if mask is not None:
            _unsupported("update_indices with a mask")
        if self.heads > 1:
            _unsupported("update_indices with heads > 1")
        x, _ = self._to_rows_layout(x)
        x = self.project_in(x)
        x = self._codebook.transform_input(x)
        self._codebook.update_indices(x, indices.reshape(x.shape[:-1]))

    update_ema_indices = update_indices

    def _loss_scratch(self, device):
        buf = getattr(self, "_loss_buf", None)
        if buf is None or buf.device != device:
            buf = torch.zeros((1,), dtype=torch.float32, device=device)
            self._loss_buf = buf
        return buf

    # ------------------------------------------------------------------ layout glue (vqp:1136-1147)
    def _to_rows_layout(self, x):
        restore = None
        if self.accept_image_fmap:
            b, c, h, w = x.shape
            x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
            restore = ("image", (h, w))
        elif self.accept_3d_fmap:
            b, c, d, h, w = x.shape
            x = x.permute(0, 2, 3, 4, 1).reshape(b, d * h * w, c)
            restore = ("3d", (d, h, w))
        elif not self.channel_last:
            x = x.transpose(1, 2)
            restore = ("transpose", None)
        return x, restore

    # ------------------------------------------------------------------ host-resident batches
    @torch.no_grad()
    def forward_host(self, x_host: torch.Tensor, n_chunks: int = 8, out=None):
        """forward() for a batch that lives in (pinned) HOST memory; results are returned in host memory.

        The batch is streamed through the GPU in `n_chunks` row chunks on three streams — upload of chunk i+1,
        kernels of chunk i and download of chunk i-1 overlap, so the call costs ~max(H2D, D2H) over PCIe instead
        of H2D + kernels + D2H.  Same arithmetic as forward(): every chunk searches the pre-update codebook, the
        chunks' EMA statistics are summed and the codebook is updated once at the end (vqp:586-617).
        `out` = optional (quantize, indices, loss) host tensors to fill (pinned for full overlap).
        The device->host copies are ASYNCHRONOUS on an internal stream that the current stream waits for: synchronise
        the current stream (or the device) before reading the returned host tensors."""
        if self.has_projections or self.accept_image_fmap or self.accept_3d_fmap or not self.channel_last or self.heads > 1:
            _unsupported("forward_host with projections / feature-map layouts / heads > 1")
        cbk = self._codebook
        emb = cbk.embed
        if not emb.is_cuda:
            raise RuntimeError("vqb200 has no CPU path: move the module to a CUDA (B200) device")
        if x_host.is_cuda or x_host.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("forward_host expects a float32 / bfloat16 CPU tensor (pinned for full overlap)")
        dev = emb.device
        shape = x_host.shape
        D = shape[-1]
        xf = x_host.reshape(-1, D)
        N = xf.shape[0]
        training = self.training
        do_update = training and not self.freeze_codebook and (cbk.ema_update or cbk.has_dead_code_replacement)
        want_loss = training and self.has_commitment_loss
        if out is None:
            out = (torch.empty(shape, dtype=x_host.dtype).pin_memory(), torch.empty(shape[:-1], dtype=torch.int64).pin_memory(),
                   torch.empty((), dtype=torch.float32).pin_memory())
        q_host, i_host, l_host = out
        qf, idf = q_host.reshape(-1, D), i_host.reshape(-1)
        bounds = host_chunk_bounds(N, n_chunks)
        n_chunks = len(bounds) - 1
        rows = max(bounds[c + 1] - bounds[c] for c in range(n_chunks))
        key = (N, D, x_host.dtype, tuple(bounds), dev)
        st = getattr(self, "_host_pipe", None)
        if st is None or st["key"] != key:
            # whole-batch device buffers (2 x 134 MB at BASELINE config 2): no upload ever waits for a buffer to be
            # recycled, so all uploads are enqueued up front and PCIe never idles on the host's enqueue pace
            st = dict(key=key, h2d=torch.cuda.Stream(dev), d2h=torch.cuda.Stream(dev),
                      x=torch.empty((N, D), dtype=x_host.dtype, device=dev),
                      q=torch.empty((N, D), dtype=x_host.dtype, device=dev),
                      i=torch.empty((N,), dtype=torch.int64, device=dev),
                      loss=torch.zeros((n_chunks,), dtype=torch.float32, device=dev),
                      stats=torch.empty((ops.stats_floats(cbk.codebook_size, D),), dtype=torch.float32, device=dev),
                      stats_chunk=torch.empty((ops.stats_floats(cbk.codebook_size, D),), dtype=torch.float32, device=dev))
            self._host_pipe = st
        cur = torch.cuda.current_stream(dev)
        up, down = st["h2d"], st["d2h"]
        up.wait_stream(cur)     # the previous call's kernels are done with x
        down.wait_stream(cur)
        ev_up = []
        with torch.cuda.stream(up):
            for c in range(n_chunks):
                r0, r1 = bounds[c], bounds[c + 1]
                st["x"][r0:r1].copy_(xf[r0:r1], non_blocking=True)
                e = torch.cuda.Event(); e.record(up); ev_up.append(e)
        weights = []
        for c in range(n_chunks):
            r0, r1 = bounds[c], bounds[c + 1]
            n = r1 - r0
            weights.append(n / N)
            cur.wait_event(ev_up[c])
            in_place = ops.STATS_MODE == 0  # fused statistics accumulate straight into the running total
            cbk.quantize_rows(st["x"][r0:r1], update=do_update, q_out=st["q"][r0:r1], idx64_out=st["i"][r0:r1],
                              loss_out=st["loss"][c:c + 1] if want_loss else None, loss_weight=self.commitment_weight,
                              stats_out=(st["stats"] if (in_place or c == 0) else st["stats_chunk"]) if do_update else None,
                              defer_ema=True, stats_accumulate=in_place and c > 0)
            if do_update and not in_place and c > 0:
                st["stats"].add_(st["stats_chunk"])
            e = torch.cuda.Event(); e.record(cur)
            with torch.cuda.stream(down):
                down.wait_event(e)
                qf[r0:r1].copy_(st["q"][r0:r1], non_blocking=True)
                idf[r0:r1].copy_(st["i"][r0:r1], non_blocking=True)
        if do_update:
            cbk.sync_stats(st["stats"])
            cbk.lerp_stats(st["stats"], normalise=cbk.ema_update and not cbk.manual_ema_update)
            if cbk.has_dead_code_replacement:   # vqp:641, over the whole batch (resident in st["x"])
                cbk.expire_codes_(cbk.transform_input(st["x"]).float())
        if want_loss:
            w = torch.tensor(weights, dtype=torch.float32, device=dev)
            loss = (st["loss"][:n_chunks] * w).sum()
            if x_host.dtype == torch.bfloat16:
                loss = loss.bfloat16().float()
            l_host.copy_(loss, non_blocking=True)
        else:
            l_host.zero_()
        cur.wait_stream(down)
        return q_host, i_host, l_host

    # ------------------------------------------------------------------ variable-length sequences (vqp:599-600, :1317-1325, :1378-1396)
    def _forward_masked(self, x, mask, freeze_codebook, ema_update, return_loss_breakdown):
        """mask (B, N) bool.  Masked positions take no part in the statistics or the loss (the reference zeroes their one-hot
        rows, vqp:599-600, and averages the loss over the unmasked elements against the ORIGINAL input, vqp:1317-1325) and come
        back as zeros / index -1.  Euclidean codebooks: the search kernel takes the mask (row_mask of vqb_vq_forward).  Cosine
        codebooks, pending k-means init, dead-code expiry: the kernels run on the compacted unmasked rows."""
        if self.has_projections or self.accept_image_fmap or self.accept_3d_fmap or not self.channel_last or self.heads > 1:
            _unsupported("mask / lens together with projections, feature-map layouts or heads > 1")
        if x.requires_grad and torch.is_grad_enabled():
            _unsupported("mask / lens on inputs that require grad")
        if not x.is_cuda:
            raise RuntimeError("vqb200 has no CPU path: inputs must live on a CUDA (B200, sm_100) device")
        assert x.ndim == 3 and mask.shape == x.shape[:2]
        freeze_codebook = self.freeze_codebook if freeze_codebook is None else freeze_codebook
        cbk = self._codebook
        ema_update = cbk.ema_update if ema_update is None else ema_update
        training = self.training
        B, N, D = x.shape
        flat = x.detach().reshape(-1, D)
        do_update = training and not freeze_codebook and (ema_update or cbk.has_dead_code_replacement)
        if (ops.STATS_MODE == 1 and not self.use_cosine_sim and cbk._initted_host and x.dtype in (torch.float32, torch.bfloat16)
                and not (do_update and cbk.has_dead_code_replacement)):
            # in-kernel mask: every row is searched (like the reference), the merge step of the search kernel drops the padding
            # rows — index -1, outputs left as pre-filled here, no loss term, no statistics — and the loss is divided by the
            # unmasked element count on the device: no host sync, no compaction pass.  (Cosine: the masked loss is taken
            # against the UN-normalised input, vqp:1319; k-means init / expiry sample from x[mask]: those take the path below.)
            flat = flat.contiguous()
            row_mask = mask.reshape(-1).contiguous().view(torch.uint8)
            n_live = row_mask.sum(dtype=torch.int64).reshape(1)
            quantize = torch.zeros_like(flat) if self.return_zeros_for_masked_padding else flat.clone()
            embed_ind = torch.full((B * N,), -1, dtype=torch.int64, device=x.device)
            want_loss = training and self.has_commitment_loss
            commit = torch.zeros((), dtype=torch.float32, device=x.device) if want_loss else None
            cbk.quantize_rows(flat, update=do_update, q_out=quantize, idx64_out=embed_ind, loss_out=commit,
                              loss_weight=self.commitment_weight, ema_update=ema_update, row_mask=row_mask, n_live=n_live)
            if want_loss:
                loss = commit.requires_grad_(torch.is_grad_enabled())
                commit_loss = commit
            else:
                loss = torch.tensor(0., device=x.device, requires_grad=training and torch.is_grad_enabled())
                commit_loss = self.zero
            quantize, embed_ind = quantize.reshape(B, N, D), embed_ind.reshape(B, N)
            if not return_loss_breakdown:
                return quantize, embed_ind, loss
            return quantize, embed_ind, loss, LossBreakdown(commit_loss, self.zero, self.zero, self.zero)
        rows = mask.reshape(-1).nonzero(as_tuple=True)[0]  # host sync (the reference's masked path syncs as well)
        quantize = torch.zeros_like(flat) if self.return_zeros_for_masked_padding else flat.clone()
        embed_ind = torch.full((B * N,), -1, dtype=torch.int64, device=x.device)
        loss = torch.tensor(0., device=x.device, requires_grad=training and torch.is_grad_enabled())
        commit_loss = self.zero
        if rows.numel() > 0:
            xc = flat[rows].contiguous()
            qc = torch.empty_like(xc)
            ic = torch.empty((xc.shape[0],), dtype=torch.int64, device=x.device)
            fused = training and self.has_commitment_loss and not self.use_cosine_sim
            commit = torch.empty((), dtype=torch.float32, device=x.device) if fused else None
            cbk.quantize_rows(xc, update=do_update, q_out=qc, idx64_out=ic, loss_out=commit,
                              loss_weight=self.commitment_weight, ema_update=ema_update)
            quantize[rows] = qc
            embed_ind[rows] = ic
            if training and self.has_commitment_loss:
                if fused:  # euclid: the original input IS what the codebook saw
                    commit_loss = commit
                    loss = commit.requires_grad_(torch.is_grad_enabled())
                else:      # cosine: mse against the un-normalised original input (vqp:1319)
                    commit_loss = F.mse_loss(qc, xc)
                    loss = loss + commit_loss * self.commitment_weight
        quantize = quantize.reshape(B, N, D)
        embed_ind = embed_ind.reshape(B, N)
        if not return_loss_breakdown:
            return quantize, embed_ind, loss
        return quantize, embed_ind, loss, LossBreakdown(commit_loss, self.zero, self.zero, self.zero)

    def _forward_separate_heads(self, x, restore, only_one, freeze_codebook, ema_update, return_loss_breakdown,
                                ema_update_weight, accum_ema_update):
        """separate_codebook_per_head (vqp:1044-1049 'b n (h d) -> h b n d', Codebook(num_codebooks=h), :1266-1268, :1354-1356):
        head i searches / updates codebook i of the (h, K, d) buffers — h independent chains on the same kernels, in head
        order (k-means init and dead-code expiry draw from the RNG head by head, like the reference's batched_sample_vectors)."""
        heads = self.heads
        b, n, hd = x.shape
        d = hd // heads
        dtype = x.dtype
        if dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"vqb200 supports float32 and bfloat16 inputs, got {dtype}")
        if accum_ema_update or ema_update_weight is not None:
            _unsupported("ema_update_weight / accum_ema_update with separate_codebook_per_head")
        input_requires_grad = x.requires_grad and torch.is_grad_enabled()
        cbk = self._codebook
        training = self.training
        do_update = training and not freeze_codebook and (ema_update or cbk.has_dead_code_replacement)
        fused_loss = training and self.has_commitment_loss and not input_requires_grad
        views = [cbk.head(i) for i in range(heads)]
        xs = [x[..., i * d:(i + 1) * d].detach().reshape(-1, d).contiguous() for i in range(heads)]
        if not cbk._initted_host:   # vqp:703: every head's k-means on the first batch, then ONE `initted` flag
            if not bool(cbk.initted):
                for v, xi in zip(views, xs):
                    v._kmeans_init(v.transform_input(xi).float())
                cbk.initted.data.copy_(torch.tensor(True))
            cbk._initted_host = True
        for v in views:
            v._initted_host = True
        loss_buf = getattr(self, "_head_loss_buf", None)
        if loss_buf is None or loss_buf.device != x.device or loss_buf.numel() != heads:
            loss_buf = self._head_loss_buf = torch.zeros((heads,), dtype=torch.float32, device=x.device)
        qs, inds = [], []
        for i, (v, xi) in enumerate(zip(views, xs)):
            q = torch.empty_like(xi)
            idx64 = torch.empty((xi.shape[0],), dtype=torch.int64, device=xi.device)
            v.quantize_rows(xi, update=do_update, q_out=q, idx64_out=idx64, loss_out=loss_buf[i:i + 1] if fused_loss else None,
                            loss_weight=self.commitment_weight, ema_update=ema_update)
            qs.append(q.reshape(b, n, d))
            inds.append(idx64.reshape(b, n))
        quantize = torch.stack(qs, dim=2)            # (b, n, h, d)
        embed_ind = torch.stack(inds, dim=-1)        # 'h b n -> b n h'  (vqp:1266-1268)
        commit_loss = self.zero
        if training and fused_loss:
            # one mse over all heads (vqp:1327) == the mean of the heads' (equal-sized) means
            commit_loss = loss_buf.mean()
            loss = commit_loss.clone().requires_grad_(torch.is_grad_enabled())
        else:
            loss = torch.tensor(0., device=x.device, requires_grad=training and torch.is_grad_enabled())  # vqp:1282
        if training:
            x_h = x.reshape(b, n, heads, d)
            if self.has_commitment_loss and not fused_loss:
                commit_loss = F.mse_loss(quantize.detach(), cbk.transform_input(x_h))
                loss = loss + commit_loss * self.commitment_weight
            if input_requires_grad and self.route_gradients_to_input:  # vqp:1225-1233
                x_t = cbk.transform_input(x_h)
                quantize = rotate_to(x_t, quantize) if self.rotation_trick else straight_through(x_t, quantize)
        quantize = self.project_out(quantize.reshape(b, n, hd))   # vqp:1354-1360
        if restore is not None:
            kind, dims = restore
            if kind == "transpose":
                quantize = quantize.transpose(1, 2)
            else:
                quantize = quantize.reshape(b, *dims, quantize.shape[-1]).movedim(-1, 1)
                embed_ind = embed_ind.reshape(b, *dims, heads)
        if only_one:
            quantize = quantize.squeeze(1)
            embed_ind = embed_ind.squeeze(1)
        if not return_loss_breakdown:
            return quantize, embed_ind, loss
        return quantize, embed_ind, loss, LossBreakdown(commit_loss if self.commitment_weight == 1. or not fused_loss else commit_loss / self.commitment_weight,
                                                       self.zero, self.zero, self.zero)

    def forward(self, x, indices=None, mask=None, lens=None, topk=None, sample_codebook_temp=None, freeze_codebook=None,
                return_loss_breakdown=False, codebook_transform_fn=None, ema_update_weight=None, accum_ema_update=False,
                ema_update=None):
        if indices is not None:
            _unsupported("forward(indices=...) cross-entropy loss")
        if mask is not None and lens is not None:
            raise AssertionError("pass either mask or lens")  # vqp:1116
        if lens is not None:  # vqp:1118-1119, :99-101
            mask = torch.arange(x.shape[1], device=lens.device) < lens[:, None]
        if mask is not None:
            return self._forward_masked(x, mask, freeze_codebook, ema_update, return_loss_breakdown)
        if topk is not None or codebook_transform_fn is not None:
            _unsupported("topk / codebook_transform_fn")
        if not x.is_cuda:
            raise RuntimeError("vqb200 has no CPU path: inputs must live on a CUDA (B200, sm_100) device")

        freeze_codebook = self.freeze_codebook if freeze_codebook is None else freeze_codebook
        ema_update = self._codebook.ema_update if ema_update is None else ema_update

        only_one = x.ndim == 2
        if only_one:
            x = x.unsqueeze(1)
        x, restore = self._to_rows_layout(x)
        x = self.project_in(x)  # vqp:1151
        heads, batch = self.heads, x.shape[0]
        if heads > 1 and self.separate_codebook_per_head:
            return self._forward_separate_heads(x, restore, only_one, freeze_codebook, ema_update, return_loss_breakdown,
                                                ema_update_weight, accum_ema_update)
        if heads > 1:  # vqp:1044-1049: 'b n (h d) -> 1 (b h) n d' — every head's sub-vector is a row for the ONE codebook
            x = x.reshape(batch, x.shape[1], heads, -1).transpose(1, 2).reshape(batch * heads, x.shape[1], -1)
        # decided AFTER project_in: with a projection the commitment loss must stay differentiable w.r.t. its weights
        # even when the raw input carries no grad (vqp:1151, :1327)
        input_requires_grad = x.requires_grad and torch.is_grad_enabled()
        shape, dtype = x.shape, x.dtype
        if dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"vqb200 supports float32 and bfloat16 inputs, got {dtype}")

        flat = x.detach().reshape(-1, shape[-1]).contiguous()
        N, D = flat.shape
        cbk = self._codebook
        training = self.training
        do_update = training and not freeze_codebook and (ema_update or cbk.has_dead_code_replacement)
        fused_loss = training and self.has_commitment_loss and not input_requires_grad

        q = torch.empty_like(flat)
        idx64 = torch.empty((N,), dtype=torch.int64, device=flat.device)
        # the kernel returns weight * mse already rounded like F.mse_loss in x.dtype (vqp:1327-1329)
        # the loss lands in a persistent scalar (stable pointer for the graph cache) and is cloned out
        loss_buf = self._loss_scratch(flat.device) if fused_loss else None
        # LossBreakdown.commitment is the UNweighted mse (vqp:1327-1329): ask the kernel for weight 1 then
        split_weight = fused_loss and return_loss_breakdown and self.commitment_weight != 1.
        cbk.quantize_rows(flat, update=do_update, q_out=q, idx64_out=idx64, loss_out=loss_buf,
                          loss_weight=1. if split_weight else self.commitment_weight, ema_update=ema_update,
                          ema_update_weight=ema_update_weight, accum_ema_update=accum_ema_update)
        commit_loss = loss_buf.clone().reshape(()) if fused_loss else self.zero
        weighted = commit_loss
        if split_weight:  # commit_loss * weight in the input dtype, promoted by the fp32 accumulator (vqp:1329, :1282)
            weighted = (commit_loss.to(dtype) * self.commitment_weight).float()

        quantize = q.reshape(shape)
        embed_ind = idx64.reshape(shape[:-1])

        if training and fused_loss:
            # vqp:1282: `loss` is a fresh fp32 scalar that requires grad in training mode
            loss = weighted.requires_grad_(torch.is_grad_enabled())
        else:
            loss = torch.tensor(0., device=flat.device, requires_grad=training and torch.is_grad_enabled())  # vqp:1282
        if training:
            if self.has_commitment_loss:
                if fused_loss:
                    pass
                else:  # differentiable w.r.t. the input: PyTorch glue on the kernel's outputs
                    x_t = cbk.transform_input(x)
                    commit_loss = F.mse_loss(quantize.detach(), x_t)
                    loss = loss + commit_loss * self.commitment_weight
            if input_requires_grad and self.route_gradients_to_input:  # vqp:1225-1233
                x_t = cbk.transform_input(x)
                quantize = rotate_to(x_t, quantize) if self.rotation_trick else straight_through(x_t, quantize)

        if heads > 1:  # vqp:1354-1358 '1 (b h) n d -> b n (h d)', :1266-1270 '1 (b h) n -> b n h'
            n = quantize.shape[1]
            quantize = quantize.reshape(batch, heads, n, -1).transpose(1, 2).reshape(batch, n, -1)
            embed_ind = embed_ind.reshape(batch, heads, n).transpose(1, 2)
        quantize = self.project_out(quantize)  # vqp:1360
        if restore is not None:  # vqp:1364-1373, :1265-1275
            kind, dims = restore
            if kind == "transpose":
                quantize = quantize.transpose(1, 2)
            else:
                b = quantize.shape[0]
                quantize = quantize.reshape(b, *dims, quantize.shape[-1]).movedim(-1, 1)
                embed_ind = embed_ind.reshape(b, *dims, *embed_ind.shape[2:])
        if only_one:
            quantize = quantize.squeeze(1)
            embed_ind = embed_ind.squeeze(1)

        if not return_loss_breakdown:
            return quantize, embed_ind, loss
        return quantize, embed_ind, loss, LossBreakdown(commit_loss, self.zero, self.zero, self.zero)
