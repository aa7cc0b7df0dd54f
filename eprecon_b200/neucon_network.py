"""NeuConNet — drop-in for models/neucon_network.py:25-624 (forward path; TSDF / occupancy refinement).

Same constructor argument (`cfg.MODEL` node), forward() signature and early-return conventions as the reference;
`outputs['coords']` (int64 [n,4]) and `outputs['tsdf']` ([n,1]) are set only when all three levels succeed.  The
coarse-to-fine loop keeps int32 coordinates on the device end to end and replaces, per level,
  upsample + Back_Project + torch.cat      -> x8 coords kernel + fused gather that writes straight into the concat buffer
  world->aligned-camera matmul + reorder   -> one kernel
  SPVCNN / GRUFusion / heads               -> eprecon_b200.modules / gru_fusion (gather-GEMM kernels)
  occupancy mask + nonzero + index         -> flags + scan compaction + row gathers
Host syncs per level: survivor count of the back-projection, voxel counts of the voxelisations, union size,
occupied count (each a 4-byte read the reference also performs).

The panoptic branch (:516-590; SURVEY.md section 8 f row 1) runs when `with_panoptic` is set (or `cfg.PANOPTIC`):
level alignment -> per-level panoptic MLPs -> SubM mask features -> eprecon_b200.mask3dformer decoder -> panoptic_post,
and `outputs['panoptic_info']` holds the reference's per-fragment list; it is None otherwise (BASELINE configs[1] is the
TSDF path).  Not built: the training losses (`loss_dict` holds zeros under the reference's keys).
"""
import numpy as np
import torch
import torch.nn as nn

try:  # the reference logs through loguru; fall back to logging so the import never fails
    from loguru import logger
except Exception:  # pragma: no cover
    import logging
    logger = logging.getLogger("eprecon_b200")

from . import _lib, executor, ops
from .gru_fusion import GRUFusion
from .mask3dformer import MultiScaleMaskedTransformerDecoder, panoptic_post
from .modules import SPVCNN, Linear4xTrans, Panoptic_Feat_Fusion
from .occupancy_initialization import Back_Project, Occupancy_Initialization
from .tensor import PointTensor

_L = _lib.lib


class NeuConNet(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.n_scales = len(cfg.THRESHOLDS) - 1
        alpha = int(self.cfg.BACKBONE2D.ARC.split("-")[-1])
        ch_in = [80 * alpha, 96 + 40 * alpha + 2, 48 + 24 * alpha + 2, 24 + 24 + 2]
        channels = [96, 48, 24]
        ch_initialization = [80, 40, 24]
        ch_initialization_down = 32
        n_views = 9
        panoptic_channels = 48
        GRU_channels = [x + y for x, y in zip(channels, ch_initialization)]
        self.channels, self.ch_img = channels, ch_initialization
        self.back_projection = nn.ModuleList()
        if self.cfg.FUSION.FUSION_ON:
            self.gru_fusion = GRUFusion(cfg, ch_in=GRU_channels, ch_voxel=channels)
            self.gru_fusion.return_int32 = True
        else:
            raise NotImplementedError("FUSION.FUSION_ON=False is not a shipped configuration (config/test.yaml:36)")
        if not self.cfg.FUSION.FULL:
            raise NotImplementedError("FUSION.FULL=False leaves grid_mask undefined in the reference (neucon_network.py:411)")
        self.sp_convs = nn.ModuleList()
        self.tsdf_preds = nn.ModuleList()
        self.occ_preds = nn.ModuleList()
        self.panoptic_preds = nn.ModuleList()
        self.initialization = Occupancy_Initialization(ch_initialization, ch_initialization_down, n_views)
        self.panoptic_feat_fusion = Panoptic_Feat_Fusion(channels[2], panoptic_channels, ch_initialization)
        # panoptic decoder, same hyper-parameters as neucon_network.py:59-71
        self.dec_layers = 6
        self.panoptic = MultiScaleMaskedTransformerDecoder(mask_classification=True, num_classes=20, hidden_dim=panoptic_channels,
                                                           num_queries=80, nheads=8, dim_feedforward=4 * panoptic_channels,
                                                           dec_layers=self.dec_layers, pre_norm=False, mask_dim=panoptic_channels)
        # Panoptic branch (neucon_network.py:516-590).  `with_panoptic_features`: level alignment, per-level panoptic MLPs and
        # the submanifold mask features only; `with_panoptic`: + decoder + post-processing -> outputs['panoptic_info'].
        # Off by default: BASELINE configs[1] (the benchmarked workload) is the TSDF path.
        self.with_panoptic_features = False
        self.with_panoptic = bool(getattr(cfg, "PANOPTIC", False))
        for i in range(len(cfg.THRESHOLDS)):
            self.back_projection.append(Back_Project(ch_initialization[i], materialize_grid=False))
            self.sp_convs.append(SPVCNN(num_classes=1, in_channels=ch_in[i], pres=1, cr=1 / 2 ** i,
                                        vres=self.cfg.VOXEL_SIZE * 2 ** (self.n_scales - i),
                                        dropout=self.cfg.SPARSEREG.DROPOUT))
            self.tsdf_preds.append(Linear4xTrans(channels[i], 1))
            self.occ_preds.append(Linear4xTrans(channels[i], 1))
            self.panoptic_preds.append(Linear4xTrans(GRU_channels[i], panoptic_channels))
        self._grid_cache = {}
        self.last_sizes = {}
        self.trace = None      # set to a dict to capture every stage's tensors (parity tests)
        self.teacher = None    # oracle trace whose data-dependent decisions (init selection, occupancy) are forced

    # ------------------------------------------------------------------------------------ reference helpers
    def upsample(self, pre_feat, pre_coords, interval, num=8):
        """API-compatible x8 upsampling (neucon_network.py:193-214); forward() uses the fused path instead."""
        with torch.no_grad():
            c32 = pre_coords.to(torch.int32).contiguous()
            up_coords = ops.upsample8(c32, interval).to(pre_coords.dtype)
            n, c = pre_feat.shape
            idx = torch.arange(n * 8, dtype=torch.int32, device=pre_feat.device)
            src = pre_feat if pre_feat.stride(1) == 1 else pre_feat.contiguous()
            up_feat = ops.gather_rows(src, c, index=idx, shift=3)[:, :c]
        return up_feat, up_coords

    def _init_grid(self, bs, interval, device):
        key = (tuple(self.cfg.N_VOX), bs, interval, str(device))
        if key not in self._grid_cache:
            axes = [torch.arange(0, n, interval, device=device, dtype=torch.int32) for n in self.cfg.N_VOX]
            g = torch.stack(torch.meshgrid(*axes, indexing="ij"), -1).view(-1, 3)
            per = [torch.cat([torch.full((g.shape[0], 1), b, dtype=torch.int32, device=device), g], 1) for b in range(bs)]
            self._grid_cache[key] = (torch.cat(per, 0).contiguous(), tuple(len(a) for a in axes))
        return self._grid_cache[key]

    def _zero_loss(self, ref):
        return torch.zeros((), dtype=torch.float32, device=ref.device)

    # ---------------------------------------------------------------------------------------------- forward
    @torch.no_grad()
    def forward(self, features, features_backbone2d_occ_pano, inputs, outputs, only_train_init=False,
                only_train_occ=False, init_overlap_count=0):
        if only_train_init or only_train_occ:
            raise NotImplementedError("training-only modes are outside the B200 inference path")
        if not self.training:
            # every BatchNorm on this path normalises with the statistics of the current fragment, as the reference does when it
            # evaluates in train() mode (main.py:357); running statistics are neither kept nor applied, so eval() is refused
            # rather than silently giving a third behaviour
            raise RuntimeError("eprecon_b200.NeuConNet runs in train() mode only (batch-statistics BatchNorm, reference main.py:357)")
        cfg = self.cfg
        bs = features[0][0].shape[0]
        dev = features[0][0].device
        loss_dict = {}
        L = _L()
        st = ops.stream_ptr()
        origin = inputs["vol_origin_partial"].float().contiguous()
        w2ac = inputs["world_to_aligned_camera"].float().contiguous()

        # ---------------------------------------------------------------- occupancy initialisation (:240-340)
        init_stage, min_view_number, occ_init_thresd = 1, 2, 0.3
        interval = 2 ** (self.n_scales - init_stage)
        scale = self.n_scales - init_stage
        up_coords, shape_init = self._init_grid(bs, interval, dev)
        KRcam = inputs["proj_matrices"][:, :, scale].permute(1, 0, 2, 3).contiguous()
        init_output = self.initialization(up_coords, origin, cfg.VOXEL_SIZE, features, KRcam, shape_init, init_stage,
                                          min_view_number)
        if init_output is None:
            loss_dict["occupancy_initialization_loss"] = self._zero_loss(origin)
            logger.warning("no valid points in initialization")
            outputs["init_overlap_count"] = init_overlap_count
            return outputs, loss_dict
        occ_init, coord_init, count_init = init_output
        src = self.initialization.last["src"]
        n_fine = up_coords.shape[0]
        fine = torch.zeros(n_fine, dtype=torch.uint8, device=dev)
        _lib.check(L.ep_scatter_selected(occ_init.data_ptr(), occ_init.stride(0), src.data_ptr(), occ_init.shape[0],
                                         occ_init_thresd, fine.data_ptr(), st), "ep_scatter_selected")
        coarse_dim = shape_init[0] // 2 ** init_stage
        assert shape_init[0] == shape_init[1] == shape_init[2], "cubic fragment volumes only"
        sel = torch.empty((bs * coarse_dim ** 3, 4), dtype=torch.int32, device=dev)
        sel_count = torch.empty(bs + 1, dtype=torch.int32, device=dev)
        _lib.check(L.ep_init_prune(fine.data_ptr(), bs, coarse_dim, interval * 2 ** init_stage, sel.data_ptr(),
                                   sel_count.data_ptr(), st), "ep_init_prune")
        n0 = int(sel_count[bs].item())
        coord_init_selected = sel[:n0]
        tr, th = self.trace, self.teacher
        if tr is not None:
            tr["init"] = {"occ": occ_init, "coords": coord_init, "count": count_init, "var": self.initialization.last["feat"]}
            tr["init_selected"] = coord_init_selected
        if th is not None:
            coord_init_selected = th["init_selected"].to(dev).to(torch.int32).contiguous()
        self.last_sizes = {"init_valid": int(occ_init.shape[0]), "n0": n0}

        # --------------------------------------------------------------------- coarse-to-fine loop (:348-511)
        pre_feat = pre_coords = None
        pre_c = 0
        pano_feats, pano_coords = [], []
        for i in range(cfg.N_LAYER):
            interval = 2 ** (self.n_scales - i)
            scale = self.n_scales - i
            if i == 0:
                up_coords = coord_init_selected.contiguous()
                min_view_number = 2
                if up_coords.shape[0] == 0:
                    loss_dict[f"tsdf_occ_loss_{i}"] = self._zero_loss(origin)
                    logger.warning("no valid points in back_projection: scale {}".format(i))
                    return outputs, loss_dict
            else:
                up_coords = ops.upsample8(pre_coords, interval)
                min_view_number = 0
            # the V per-view maps of this pyramid level go straight into the channels-last gather buffer (one launch; zero-copy
            # when the caller hands channels-last storage) -- the reference stacks them (neucon_network.py:364)
            feats_nhwc = ops.pack_views_nhwc([feat[scale] for feat in features_backbone2d_occ_pano])
            KRcam = inputs["proj_matrices"][:, :, scale].permute(1, 0, 2, 3).contiguous().float()
            c_img = feats_nhwc.shape[4]
            c_cat = c_img + pre_c
            res = ops.backproject(up_coords, origin, cfg.VOXEL_SIZE, feats_nhwc, KRcam, min_view_number,
                                  mode="mean", want_src=True, alloc_width=ops.ceil4(c_cat))
            if res is None:
                loss_dict[f"tsdf_occ_loss_{i}"] = self._zero_loss(origin)
                logger.warning("no valid points in back_projection: scale {}".format(i))
                return outputs, loss_dict
            up_coords = res["coords"]
            feat = res["buffer"]                                  # [M, ceil4(c_img + pre_c)]; cols [0,c_img) = volume
            m = feat.shape[0]
            if ops.ceil4(c_cat) != c_cat:
                feat[:, c_cat:].zero_()
            if i != 0:  # concat the parent's features: up_feat[count >= min] == pre_feat[src >> 3]
                ops.gather_rows(pre_feat, pre_c, index=res["src"], shift=3, out=feat, out_col=c_img)
            r_coords = ops.aligned_coords(up_coords, origin, cfg.VOXEL_SIZE, w2ac)
            feat_v = self.sp_convs[i](PointTensor(feat, r_coords))        # [M, channels[i]]
            cv = self.channels[i]
            if tr is not None:
                tr[f"l{i}_pre_gru"] = {"coords": up_coords, "feat_in": feat[:, :c_cat], "pts": r_coords, "spvcnn": feat_v,
                                       "count": res["count"]}
            feat_all = torch.empty((m, cv + c_img), dtype=torch.float32, device=dev)
            feat_all[:, :cv] = feat_v
            feat_all[:, cv:] = feat[:, :c_img]
            up_coords, feat_all, tsdf_target, occ_target = self.gru_fusion(up_coords, feat_all, inputs, i)
            feat_v = feat_all[:, :cv]
            if executor.enabled() and feat_v.stride(0) % 4 == 0:
                tsdf, occ = executor.linear4x(self, [self.tsdf_preds[i], self.occ_preds[i]], feat_v)   # both heads, one call
            else:
                tsdf = self.tsdf_preds[i](feat_v)
                occ = self.occ_preds[i](feat_v)
            loss_dict[f"tsdf_occ_loss_{i}"] = self._zero_loss(origin)

            # ------------------------------------------------ sparsity for the next level (:454-507)
            flags = ops.threshold_flags(occ, float(cfg.THRESHOLDS[i]), mode=0)
            u = up_coords.shape[0]
            if tr is not None:
                tr[f"l{i}"] = {"coords": up_coords, "feat_all": feat_all, "tsdf": tsdf, "occ": occ,
                               "occupancy": flags.bool().clone(), "occ_target": occ_target}
            if th is not None:
                flags = th[f"l{i}"]["occupancy"].to(dev).to(torch.uint8).contiguous()
            exceed_num = 1.5
            if bs == 1:
                index, num = ops.compact_flags(flags)
                nums = [num]
            else:
                nums = [int(flags[up_coords[:, 0] == b].sum().item()) for b in range(bs)]
                index = None
            for b, num_batch in enumerate(nums):
                if num_batch < 500:
                    logger.warning("no valid points: scale {}".format(i))
                    return outputs, loss_dict
                if self.training and num_batch > cfg.TRAIN_NUM_SAMPLE[i] * exceed_num:
                    logger.warning("exceed too many points: scale {} num_batch {}".format(i, num_batch))
                    return outputs, loss_dict
                elif self.training and num_batch > cfg.TRAIN_NUM_SAMPLE[i]:
                    logger.warning("choice too many points: scale {} num_batch {}".format(i, num_batch))
                    choice = np.random.choice(num_batch, num_batch - cfg.TRAIN_NUM_SAMPLE[i], replace=False)
                    rows = torch.nonzero((up_coords[:, 0] == b) & flags.bool()).squeeze(1)
                    flags[rows[torch.from_numpy(choice).to(dev)]] = 0
                    index = None
            if index is None:
                index, num = ops.compact_flags(flags)
            if occ_target is not None:
                ot = occ_target.view(-1)
                for b in range(bs):
                    hit = ot[index.long()] if bs == 1 else ot[index.long()][up_coords[index.long(), 0] == b]
                    if int(hit.sum().item()) == 0:
                        logger.warning("occ_target is 0: scale {} num_batch {}".format(i, 0))
                        return outputs, loss_dict
            pre_coords = ops.gather_coords(up_coords, index)
            pre_c = cv + 2
            pre_feat = torch.empty((num, ops.ceil4(pre_c)), dtype=torch.float32, device=dev)
            if ops.ceil4(pre_c) != pre_c:
                pre_feat[:, pre_c:].zero_()
            ops.gather_rows(feat_all, cv, index=index, out=pre_feat, out_col=0)
            tsdf_c = tsdf if tsdf.stride(1) == 1 else tsdf.contiguous()
            occ_c = occ if occ.stride(1) == 1 else occ.contiguous()
            pre_tsdf = ops.gather_rows(tsdf_c, 1, index=index)
            pre_feat[:, cv] = pre_tsdf[:, 0]
            pre_feat[:, cv + 1] = ops.gather_rows(occ_c, 1, index=index)[:, 0]
            if self.with_panoptic_features or self.with_panoptic:
                pano_feats.append(ops.gather_rows(feat_all, cv + c_img, index=index))
                pano_coords.append(pre_coords)
            self.last_sizes[f"level{i}"] = {"candidates": int(res["count"].shape[0]), "projected": m, "fused": u,
                                            "occupied": num}
            if i == cfg.N_LAYER - 1:
                outputs["coords"] = pre_coords.long()
                outputs["tsdf"] = pre_tsdf[:, :1]
        outputs["panoptic_info"] = None
        if self.with_panoptic_features or self.with_panoptic:
            outputs["panoptic_features"] = self.panoptic_prepare(pano_coords, pano_feats, bs)
            if self.with_panoptic:
                outputs["panoptic_outs"], outputs["panoptic_info"] = self.panoptic_decode(outputs["panoptic_features"], bs)
        return outputs, loss_dict

    @torch.no_grad()
    def panoptic_decode(self, pf, bs):
        """models/neucon_network.py:560-586: per fragment, decoder over the three aligned levels + panoptic_post."""
        shape = tuple(int(n) for n in self.cfg.N_VOX)
        outs, infos = [], []
        for b in range(bs):
            sel = [slice(None) if bs == 1 else torch.nonzero(pf["coords"][p][:, 0] == b).squeeze(1) for p in range(3)]
            feats = [pf["feats"][p][sel[p]][:, :48].unsqueeze(0).permute(0, 2, 1) for p in range(3)]
            coords = [pf["coords"][p][sel[p]][:, 1:].unsqueeze(0) for p in range(3)]
            mask = pf["mask_features"][sel[2]][:, :48].unsqueeze(0).permute(0, 2, 1)
            out = self.panoptic(panoptic_features=feats, panoptic_coords=coords, mask_features=mask, spitial_shape=shape)
            outs.append(out)
            infos.append(panoptic_post(out))
        return outs, infos

    @torch.no_grad()
    def panoptic_prepare(self, coords, feats, bs):
        """models/neucon_network.py:516-560: (1) keep a level-1 / level-0 voxel only if an occupied level-2 voxel lies inside
        it (the reference's O(N*M) row compare becomes parent marking in a byte volume), (2) per-level Linear4xTrans to 48
        channels, (3) three residual submanifold convs on the level-2 sites -> mask features."""
        L = _L()
        st = ops.stream_ptr()
        dev = coords[2].device
        dims = [int(n) for n in self.cfg.N_VOX]
        out_c, out_f = [None, None, coords[2]], [None, None, feats[2]]
        for lvl, step in ((1, 2), (0, 4)):
            d = [n // step for n in dims]
            vol = torch.zeros(bs * d[0] * d[1] * d[2], dtype=torch.uint8, device=dev)
            _lib.check(L.ep_mark_parents(coords[2].data_ptr(), coords[2].shape[0], step, d[0], d[1], d[2], bs, vol.data_ptr(), st),
                       "ep_mark_parents")
            n = coords[lvl].shape[0]
            flags = torch.empty(n, dtype=torch.uint8, device=dev)
            _lib.check(L.ep_lookup_marks(coords[lvl].data_ptr(), n, step, d[0], d[1], d[2], bs, vol.data_ptr(), flags.data_ptr(),
                                         st), "ep_lookup_marks")
            index, _ = ops.compact_flags(flags)
            out_c[lvl] = ops.gather_coords(coords[lvl], index)
            out_f[lvl] = ops.gather_rows(feats[lvl], feats[lvl].shape[1], index=index)
        pf = [self.panoptic_preds[p](out_f[p]) for p in range(3)]
        shape = tuple(dims)
        mask_chunks = []
        for b in range(bs):
            sel = slice(None) if bs == 1 else torch.nonzero(out_c[2][:, 0] == b).squeeze(1)
            cb = out_c[2][sel]
            mask_chunks.append(self.panoptic_feat_fusion.generate_mask_features(
                panoptic_feats=pf[2][sel], coords_b=torch.zeros_like(cb[:, 0]), coords_xyz=cb[:, 1:], batch_size=1,
                spitial_shape=shape))
        return {"coords": [c.long() for c in out_c], "feats": pf,
                "mask_features": mask_chunks[0] if bs == 1 else torch.cat(mask_chunks, 0)}
