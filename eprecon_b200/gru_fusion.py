"""GRUFusion — drop-in for models/gru_fusion.py:8-394 (feature fusion + direct-substitute TSDF fusion).

Same constructor, forward() signature, return structure and per-scene state as the reference.  What changes
is the mechanism: the reference densifies the current fragment and the cropped global state into two
[d,d,d,C] fp32 volumes per level (9.7 / 38.9 / 169.9 MB each) just to take their union and re-gather; here the
union is a compaction over two int32 row-index volumes [d^3] (csrc/union_gather.cu) followed by two row
gathers, and the ConvGRUs of a level share one set of voxelisations / kernel maps.

Ground-truth bookkeeping (target_tsdf_volume; only used for the reference's loss and its "GT occupancy overlaps
prediction" guard, models/neucon_network.py:486-490) stays a small dense 1-channel volume in plain torch.
Panoptic global fusion (direct-substitute mode with `panoptic_infos`; gru_fusion.py:116-193,352-369; SURVEY.md
section 8 f2): the reference computes the IoU of every (fragment segment, scene instance) candidate pair with an
M x N pairwise distance matrix (`compute_overlap`) inside a host loop with `.item()` reads.  Here the union kernel
already knows, for every union site, the row of the scene voxel at the same coordinate (`row_b`), so all pair
intersections are ONE bincount over (segment id, scene instance id) and the host loop runs on three small tables
read back once.
"""
import torch
import torch.nn as nn

from . import _lib, executor, ops, sparse
from .modules import ConvGRU
from .ops import stream_ptr

_L = _lib.lib


class GRUFusion(nn.Module):
    def __init__(self, cfg, ch_in=None, direct_substitute=False, trianing=True, ch_voxel=None):
        super().__init__()
        self.cfg = cfg
        self.direct_substitude = direct_substitute
        self.trianing = trianing
        if direct_substitute:
            self.ch_in, self.feat_init = [1, 1, 1], 1
        else:
            self.ch_in, self.feat_init = ch_in, 0
        if ch_voxel is not None:
            self.ch_voxel = ch_voxel
            self.ch_img = [x - y for x, y in zip(ch_in, ch_voxel)]
        self.n_scales = len(cfg.THRESHOLDS) - 1
        self.scene_name = [None, None, None]
        self.global_origin = [None, None, None]
        self.global_volume = [None, None, None]      # dict(C int32 [Ng,3] global voxel ids, F fp32 [Ng,C])
        self.target_tsdf_volume = [None, None, None]
        self.global_instance = None                   # int32 [Ng] scene instance / semantic id per row of the finest
        self.global_semantic = None                   # global volume (direct substitute + panoptic_infos only)
        self.return_int32 = False
        self._consts = {}
        if direct_substitute:
            self.fusion_nets_voxel = self.fusion_nets_img = None
        else:
            self.fusion_nets_voxel = nn.ModuleList()
            self.fusion_nets_img = nn.ModuleList()
            for i, ch in enumerate(self.ch_voxel):
                self.fusion_nets_voxel.append(ConvGRU(hidden_dim=ch, input_dim=ch, pres=1,
                                                      vres=self.cfg.VOXEL_SIZE * 2 ** (self.n_scales - i)))
            for i, ch in enumerate(self.ch_img):
                self.fusion_nets_img.append(ConvGRU(hidden_dim=ch, input_dim=ch, pres=1,
                                                    vres=self.cfg.VOXEL_SIZE * 2 ** (self.n_scales - i)))

    def _const(self, values, device, dtype):
        """Small constant device tensors (volume dims, relative origins) cached per value: a torch.tensor(list) H2D copy
        is a synchronising 200 us detour on the hot path."""
        key = (values, str(device), dtype)
        t = self._consts.get(key)
        if t is None:
            if len(self._consts) > 256:
                self._consts.clear()
            t = self._consts[key] = torch.tensor(values, dtype=dtype, device=device)
        return t

    def reset(self, i, device="cuda"):
        c = self.ch_in[i]
        self.global_volume[i] = {"C": torch.zeros((0, 3), dtype=torch.int32, device=device),
                                 "F": torch.zeros((0, ops.ceil4(c)), dtype=torch.float32, device=device)}
        self.target_tsdf_volume[i] = {"C": torch.zeros((0, 3), dtype=torch.int64, device=device),
                                      "F": torch.zeros((0, 1), dtype=torch.float32, device=device)}
        self.global_instance = torch.zeros(0, dtype=torch.int32, device=device)
        self.global_semantic = torch.zeros(0, dtype=torch.int32, device=device)

    # ------------------------------------------------------------------------------------------------
    def panoptic_fusion(self, scale, global_valid, relative_origin, panoptic_info, current_coords, row_b=None):
        """gru_fusion.py:133-193 (+ compute_overlap :116-131).  panoptic_info['panoptic_seg'] = [segment id per union
        site int32 [U], segments_info]; `row_b` int32 [U] = row of the scene voxel at the same coordinate (-1: none).
        Returns (new_current_instance, new_current_semantic) int32 [U].  A thing segment takes the id of the first
        (ascending) scene instance of its class whose IoU with it exceeds 0.05, else a fresh id; stuff takes its class id."""
        seg, info = panoptic_info["panoptic_seg"]
        dev = seg.device
        n_seg = len(info)
        gi, gs = self.global_instance, self.global_semantic
        stuff_max, thr = 2, 0.05
        seg_l = seg.long()
        if gi.shape[0] > 0 and n_seg > 0:
            top = torch.stack([gi.max(), gs.max()]).tolist()                      # host read 1: table extents
            m_ins, m_sem = int(top[0]) + 1, int(top[1]) + 1
            hit = row_b >= 0
            pair = torch.bincount(seg_l[hit] * m_ins + gi[row_b[hit].long()].long(), minlength=(n_seg + 1) * m_ins)
            area_seg = torch.bincount(seg_l, minlength=n_seg + 1)[: n_seg + 1]
            area_ins = torch.bincount(gi.long(), minlength=m_ins)
            gv = global_valid.nonzero().squeeze(1)
            present = torch.bincount(gs[gv].long() * m_ins + gi[gv].long(), minlength=m_sem * m_ins)
            host = torch.cat([pair, area_seg, area_ins, present]).tolist()        # host read 2: all pair statistics
            o = 0
            pair_h = host[o:o + (n_seg + 1) * m_ins]; o += (n_seg + 1) * m_ins    # noqa: E702
            aseg_h = host[o:o + n_seg + 1]; o += n_seg + 1                        # noqa: E702
            ains_h = host[o:o + m_ins]; o += m_ins                                # noqa: E702
            pres_h = host[o:o + m_sem * m_ins]
            max_id = max(m_ins - 1, stuff_max)
        else:
            m_ins = m_sem = 0
            pair_h = aseg_h = ains_h = pres_h = []
            max_id = max(int(gi.max()), stuff_max) if gi.shape[0] > 0 else stuff_max
        tab_i, tab_s = [0] * (n_seg + 1), [0] * (n_seg + 1)
        inc = 1
        f32 = lambda v: torch.tensor(float(v), dtype=torch.float32)  # noqa: E731   (the reference divides int64 tensors)
        for i, d in enumerate(info):
            cls = int(d["category_id"])
            if not d["isthing"]:
                tab_i[i + 1] = tab_s[i + 1] = cls
                continue
            matched = False
            if cls < m_sem:
                for ins in range(m_ins):                                          # ascending, as torch.unique
                    if not pres_h[cls * m_ins + ins]:
                        continue
                    inter = pair_h[(i + 1) * m_ins + ins]
                    union = aseg_h[i + 1] + ains_h[ins] - inter
                    if union > 0 and bool(f32(inter) / f32(union) > thr):
                        tab_i[i + 1], tab_s[i + 1] = ins, cls
                        matched = True
                        break
            if not matched:
                tab_i[i + 1], tab_s[i + 1] = max_id + inc, cls
                inc += 1
        ti = torch.tensor(tab_i, dtype=torch.int32, device=dev)
        tsm = torch.tensor(tab_s, dtype=torch.int32, device=dev)
        return ti[seg_l], tsm[seg_l]

    # ------------------------------------------------------------------------------------------------
    def _union(self, coords_b4, values, c, rel, dims, scale, batch, interval):
        """Sparse union of the current rows and the in-volume global rows.  Returns level-unit coords (b,x,y,z)
        int32 [U,4], row ids into current / global (-1 = absent) and the global `valid` mask."""
        dev = values.device
        L = _L()
        st = stream_ptr()
        dx, dy, dz = dims
        nvol = dx * dy * dz
        g = self.global_volume[scale]
        mode = 1 if self.direct_substitude else 0
        vol_a = torch.empty(nvol, dtype=torch.int32, device=dev)
        vol_b = torch.empty(nvol, dtype=torch.int32, device=dev)
        _lib.check(L.ep_fill_i32(vol_a.data_ptr(), nvol, -1, st), "ep_fill_i32")
        _lib.check(L.ep_fill_i32(vol_b.data_ptr(), nvol, -1, st), "ep_fill_i32")
        n = coords_b4.shape[0]
        _lib.check(L.ep_scatter_rows_to_volume(coords_b4.data_ptr(), 4, 1, n, interval, 0, 0, 0, dx, dy, dz,
                                               values.data_ptr(), values.stride(0), c, mode, vol_a.data_ptr(), 0, st),
                   "ep_scatter_rows_to_volume")
        ng = g["C"].shape[0]
        valid = torch.zeros(ng, dtype=torch.bool, device=dev)
        if ng > 0:
            _lib.check(L.ep_scatter_rows_to_volume(g["C"].data_ptr(), 3, 0, ng, 1, rel[0], rel[1], rel[2], dx, dy, dz,
                                                   g["F"].data_ptr(), g["F"].stride(0), c, mode, vol_b.data_ptr(),
                                                   valid.data_ptr(), st), "ep_scatter_rows_to_volume")
        flags = torch.empty(nvol, dtype=torch.uint8, device=dev)
        _lib.check(L.ep_union_flags(vol_a.data_ptr(), vol_b.data_ptr(), nvol, flags.data_ptr(), st), "ep_union_flags")
        sites, u = ops.compact_flags(flags)
        out_coords = torch.empty((u, 4), dtype=torch.int32, device=dev)
        row_a = torch.empty(u, dtype=torch.int32, device=dev)
        row_b = torch.empty(u, dtype=torch.int32, device=dev)
        if u > 0:
            _lib.check(L.ep_union_sites(sites.data_ptr(), u, dy, dz, batch, 1, vol_a.data_ptr(), vol_b.data_ptr(),
                                        out_coords.data_ptr(), row_a.data_ptr(), row_b.data_ptr(), st), "ep_union_sites")
        return out_coords, row_a, row_b, valid

    def _fuse_targets(self, inputs, i, scale, rel, dims, updated_xyz):
        """Reference GT fusion (gru_fusion.py:98-112,205-215,324-327) on a dense 1-channel volume.  Same result as the
        reference's nonzero / boolean-index formulation, written with masked dense ops so that the only host round-trips
        are the two data-dependent list sizes of the updated sparse GT map (one when the scene has no GT state yet)."""
        lvl = self.cfg.N_LAYER - scale - 1
        occ_t = inputs["occ_list"][lvl][i]
        tsdf_full = inputs["tsdf_list"][lvl][i]
        tgt = self.target_tsdf_volume[scale]
        dev = occ_t.device
        relt = self._const(tuple(rel), dev, torch.int64)
        nvol = dims[0] * dims[1] * dims[2]
        flat = torch.full((nvol + 1,), 1.0, dtype=tgt["F"].dtype, device=dev)     # +1: dump slot for out-of-volume rows
        ng = tgt["C"].shape[0]
        valid_t = None
        if ng > 0:
            dim = self._const(tuple(dims), dev, torch.int64)
            gc = tgt["C"] - relt
            valid_t = ((gc < dim) & (gc >= 0)).all(dim=-1)
            lin = (gc[:, 0] * dims[1] + gc[:, 1]) * dims[2] + gc[:, 2]
            flat[torch.where(valid_t, lin, torch.full_like(lin, nvol))] = tgt["F"][:, 0]
        vol = flat[:nvol].view(dims[0], dims[1], dims[2])
        vol = torch.where(occ_t, tsdf_full.to(vol.dtype), vol)                    # the current fragment's GT wins
        u = updated_xyz.long()
        tsdf_target = vol[u[:, 0], u[:, 1], u[:, 2]].unsqueeze(-1)
        occ_target = tsdf_target.abs() < 1
        # update the GT map: drop the in-volume rows, append every |tsdf| < 1 voxel of the fused local volume
        keep = torch.nonzero(vol.abs() < 1)
        new_f = vol[keep[:, 0], keep[:, 1], keep[:, 2]].unsqueeze(-1)
        new_c = keep + relt
        if ng > 0:
            stay = torch.nonzero(valid_t == False).squeeze(1)  # noqa: E712
            tgt["F"] = torch.cat([tgt["F"][stay], new_f])
            tgt["C"] = torch.cat([tgt["C"][stay], new_c])
        else:
            tgt["F"], tgt["C"] = new_f, new_c
        return tsdf_target, occ_target

    def save_mesh(self, scale, outputs, scene):
        """Scene TSDF as a dense bounding-box volume (gru_fusion.py:217-257), TSDF channel only."""
        if outputs is None:
            outputs = dict()
        keys = ("origin", "scene_tsdf", "scene_name", "scene_instance", "scene_semantic")
        if "scene_name" not in outputs:
            for k in keys:
                outputs[k] = []
        if scene in outputs["scene_name"]:
            idx = outputs["scene_name"].index(scene)
            for k in keys:
                del outputs[k][idx]
        outputs["scene_name"].append(scene)
        g = self.global_volume[scale]
        c = g["C"].long()
        max_c, min_c = c.max(0)[0], c.min(0)[0]
        outputs["origin"].append(min_c * self.cfg.VOXEL_SIZE * (2 ** (self.cfg.N_LAYER - scale - 1)))
        ind = c - min_c
        dim = (max_c - min_c + 1).tolist()
        vol = torch.full(dim, 1.0, dtype=torch.float32, device=c.device)
        vol[ind[:, 0], ind[:, 1], ind[:, 2]] = g["F"][:, 0]
        outputs["scene_tsdf"].append(vol)
        for key, ids in (("scene_instance", self.global_instance), ("scene_semantic", self.global_semantic)):
            v = torch.zeros(dim, dtype=torch.int32, device=c.device)
            if ids is not None and ids.shape[0] == c.shape[0]:          # fused with panoptic_infos (gru_fusion.py:251-256)
                v[ind[:, 0], ind[:, 1], ind[:, 2]] = ids
            outputs[key].append(v)
        return outputs

    @torch.no_grad()
    def forward(self, coords, values_in, inputs, scale=2, outputs=None, save_mesh=False, panoptic_infos=None):
        batch_size = len(inputs["fragment"])
        interval = 2 ** (self.cfg.N_LAYER - scale - 1)
        c_all = self.ch_in[scale]
        dev = values_in.device
        coords32 = coords.to(torch.int32).contiguous()
        values_in = values_in if (values_in.stride(1) == 1 and values_in.stride(0) % 4 == 0) else values_in.contiguous()
        if values_in.stride(0) % 4 != 0:  # e.g. the [N,1] TSDF column
            tmp = torch.zeros((values_in.shape[0], ops.ceil4(c_all)), dtype=torch.float32, device=dev)
            tmp[:, :c_all] = values_in
            values_in = tmp
        coords_all, values_all, tsdf_all, occ_all = [], [], [], []
        dims = [int(n) // 2 ** (self.cfg.N_LAYER - scale - 1) for n in self.cfg.N_VOX]
        for i in range(batch_size):
            scene = inputs["scene"][i]
            global_origin = inputs["vol_origin"][i]
            origin = inputs["vol_origin_partial"][i]
            if scene != self.scene_name[scale] and self.scene_name[scale] is not None and self.direct_substitude:
                outputs = self.save_mesh(scale, outputs, self.scene_name[scale])
            if self.scene_name[scale] is None or scene != self.scene_name[scale]:
                self.scene_name[scale] = scene
                self.reset(scale, dev)
                self.global_origin[scale] = global_origin
            voxel_size = self.cfg.VOXEL_SIZE * interval
            rel = ((origin - self.global_origin[scale]) / voxel_size).long().tolist()   # trunc, as .long() does
            if batch_size == 1:
                cb, vb = coords32, values_in
            else:
                sel = torch.nonzero(coords32[:, 0] == i).squeeze(1)
                if sel.numel() == 0:
                    continue
                cb, vb = coords32[sel].contiguous(), values_in[sel].contiguous()
            if cb.shape[0] == 0:
                continue
            upd, row_a, row_b, valid = self._union(cb, vb, c_all, rel, dims, scale, i, interval)
            u = upd.shape[0]
            g = self.global_volume[scale]
            values = ops.gather_rows(vb, c_all, index=row_a, fill=float(self.feat_init), m=u)
            gvalues = ops.gather_rows(g["F"], c_all, index=row_b, fill=float(self.feat_init), m=u) if g["F"].shape[0] \
                else torch.full((u, ops.ceil4(c_all)), float(self.feat_init), dtype=torch.float32, device=dev)
            if "occ_list" in inputs:
                tsdf_target, occ_target = self._fuse_targets(inputs, i, scale, rel, dims, upd[:, 1:])
            else:
                tsdf_target = occ_target = None
            if not self.direct_substitude:
                cv = self.ch_voxel[scale]
                org = origin.float().view(1, 3).contiguous()
                w2ac = inputs["world_to_aligned_camera"][i].float().view(1, 4, 4).contiguous()
                upd0 = upd.clone()
                upd0[:, 0] = 0
                r_coords = ops.aligned_coords(upd0, org, voxel_size, w2ac, zero_batch=True)
                gru_v, gru_i = self.fusion_nets_voxel[scale], self.fusion_nets_img[scale]
                if executor.enabled() and c_all % 4 == 0 and cv % 4 == 0:
                    # both ConvGRUs of the level on shared voxelisations as ONE native call (csrc/executor.cu)
                    values = executor.gru_level(self, gru_v, gru_i, r_coords, gvalues, values, cv, c_all)
                else:
                    pc1 = sparse.PointCloud(r_coords, gru_v.vres)
                    pc2 = sparse.PointCloud(pc1.scaled, gru_v.vres, order="hash")
                    out_v = gru_v.run(gvalues[:, :cv], values[:, :cv], pc1, pc2)
                    out_i = gru_i.run(gvalues[:, cv:c_all], values[:, cv:c_all], pc1, pc2)
                    values = torch.cat([out_v[:, :cv], out_i[:, :c_all - cv]], dim=-1)
            new_inst = new_sem = None
            if self.direct_substitude and panoptic_infos is not None:
                # gru_fusion.py:352-363: the fragment's segment ids at the union sites (0 where only the scene has a voxel)
                info = panoptic_infos[i]
                # panoptic_seg is already per fragment (NeuConNet.panoptic_decode emits one per batch entry over that entry's
                # voxels, rows aligned with this entry's coords; the reference pairs it directly, gru_fusion.py:355)
                seg = info["panoptic_seg"][0].to(torch.int32)
                info["panoptic_seg"][0] = torch.where(row_a >= 0, seg[row_a.clamp_min(0).long()], torch.zeros((), dtype=torch.int32, device=dev))
                new_inst, new_sem = self.panoptic_fusion(scale=scale, global_valid=valid, relative_origin=rel, panoptic_info=info,
                                                         current_coords=upd[:, 1:], row_b=row_b)
            # update_map (gru_fusion.py:195-204): drop in-volume global rows, append the fused ones
            relt = self._const(tuple(rel), dev, torch.int32)
            vpad = values if values.shape[1] == g["F"].shape[1] else torch.nn.functional.pad(
                values, (0, g["F"].shape[1] - values.shape[1]))
            if g["C"].shape[0] > 0:
                stay = torch.nonzero(valid == False).squeeze(1)  # noqa: E712  (one size read-back for both gathers)
                g["F"] = torch.cat([g["F"][stay], vpad])
                g["C"] = torch.cat([g["C"][stay], upd[:, 1:] + relt])
                if new_inst is not None:
                    self.global_instance = torch.cat([self.global_instance[stay], new_inst])
                    self.global_semantic = torch.cat([self.global_semantic[stay], new_sem])
            else:
                g["F"], g["C"] = vpad, upd[:, 1:] + relt
                if new_inst is not None:
                    self.global_instance, self.global_semantic = new_inst, new_sem
            out_c = upd.clone()
            out_c[:, 1:] *= interval
            coords_all.append(out_c)
            values_all.append(values[:, :c_all])
            if tsdf_target is not None:
                tsdf_all.append(tsdf_target)
                occ_all.append(occ_target)
            if self.direct_substitude and save_mesh:
                outputs = self.save_mesh(scale, outputs, self.scene_name[scale])
        if self.direct_substitude:
            return outputs
        if not coords_all:
            return None, None, None, None
        def cat(xs):
            return xs[0] if len(xs) == 1 else torch.cat(xs)

        coords_out = cat(coords_all)
        if not self.return_int32:
            coords_out = coords_out.long()
        return (coords_out, cat(values_all), cat(tsdf_all) if tsdf_all else None, cat(occ_all) if occ_all else None)
