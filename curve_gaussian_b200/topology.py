"""Curve-set surgery and optimizer bookkeeping of the reference GaussianCurveModel.

These are the calls train.py makes every 500-2000 iterations between hot-path steps
(scene/gaussian_curve_model.py:200-456; scene/gaussian_model.py:460-533 for the optimizer edits). They only
rearrange the per-curve tensors `_curve_points/_width/_opacity/_mask/is_bezier` (+ Adam moments); everything
runs as a handful of whole-tensor device ops, no per-curve host loop, and ends with `prepare_scaling_rot()` so
the hot path sees the new curve set. Method names, arguments and results follow the reference.

`merge_curves` is carried over in a deterministic form: the reference draws a RANSAC consensus set before each
line / Bezier refit (skimage, unseeded); here every sampled point of an already pre-filtered pair takes part, the
fits are closed-form and batched (see the method).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch
from torch import nn


def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """Log-linear learning-rate decay with an optional eased-in start (utils/general_utils.py:99-133)."""
    log_i = math.log(lr_init) if lr_init > 0 else 0.0
    log_f = math.log(lr_final) if lr_final > 0 else 0.0

    def helper(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        rate = 1.0
        if lr_delay_steps > 0:
            rate = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0), 1))
        t = min(max(step / max_steps, 0), 1)
        return rate * math.exp(log_i * (1 - t) + log_f * t)

    return helper


def de_casteljau_split(curves: torch.Tensor, t: torch.Tensor, is_bezier: torch.Tensor):
    """Split every curve at its own parameter t into (left, right) control polygons, each (K,4,3).

    Cubic Beziers by De Casteljau; straight segments (is_bezier False) are cut at the point (1-t)P0 + tP3 and
    both halves get evenly spaced inner control points (gaussian_curve_model.py:389-424). t is (K,1).
    """
    p0, p1, p2, p3 = curves.unbind(1)
    s = 1 - t
    q0, q1, q2 = s * p0 + t * p1, s * p1 + t * p2, s * p2 + t * p3
    r0, r1 = s * q0 + t * q1, s * q1 + t * q2
    mid = s * r0 + t * r1
    left = torch.stack([p0, q0, r0, mid], dim=1)
    right = torch.stack([mid, r1, q2, p3], dim=1)
    if bool(is_bezier.all()):
        return left, right
    cut = s * p0 + t * p3
    left_s = torch.stack([p0, (2 / 3) * p0 + (1 / 3) * cut, (1 / 3) * p0 + (2 / 3) * cut, cut], dim=1)
    right_s = torch.stack([cut, (2 / 3) * cut + (1 / 3) * p3, (1 / 3) * cut + (2 / 3) * p3, p3], dim=1)
    sel = is_bezier[:, None, None]
    return torch.where(sel, left, left_s), torch.where(sel, right, right_s)


def de_casteljau_trim(curves, from_t, end_t, is_bezier):
    """Keep the part after from_t, then the part of THAT curve before end_t (gaussian_curve_model.py:367-370;
    end_t is applied in the re-parametrised remainder, as the reference does)."""
    _, rest = de_casteljau_split(curves, from_t, is_bezier)
    kept, _ = de_casteljau_split(rest, end_t, is_bezier)
    return kept


def resample_mask_rows(mask: torch.Tensor, start: torch.Tensor, end: torch.Tensor) -> torch.Tensor:
    """Row b of the (B,n) mask, restricted to [start_b, end_b], stretched back to n samples with the
    half-pixel-centre linear rule of `F.interpolate(mode='bilinear', align_corners=False)` - all rows at once
    (the reference loops over the curves, gaussian_curve_model.py:446-450)."""
    B, n = mask.shape
    length = (end - start + 1).to(mask.dtype)                      # source samples per row
    i = torch.arange(n, device=mask.device, dtype=mask.dtype)[None, :]
    src = ((i + 0.5) * (length[:, None] / n) - 0.5).clamp_min(0.0)
    lo = src.floor().long()
    lo = torch.minimum(lo, (end - start)[:, None])
    hi = torch.minimum(lo + 1, (end - start)[:, None])
    w = src - lo.to(mask.dtype)
    a = torch.gather(mask, 1, start[:, None] + lo)
    b = torch.gather(mask, 1, start[:, None] + hi)
    return (1 - w) * a + w * b


# ----------------------------------------------------------------------------------------------------------
# Adam surgery: every param group holds exactly one tensor (gaussian_curve_model.py:203-210)
def _edit_groups(optimizer, names, new_param: Callable[[str, torch.Tensor], torch.Tensor],
                 new_moment: Callable[[str, torch.Tensor, torch.Tensor], torch.Tensor]) -> Dict[str, nn.Parameter]:
    out = {}
    for group in optimizer.param_groups:
        name = group["name"]
        if names is not None and name not in names:
            continue
        old = group["params"][0]
        state = optimizer.state.pop(old, None)
        new = nn.Parameter(new_param(name, old.detach()).requires_grad_(True))
        if state is not None and "exp_avg" in state:
            state["exp_avg"] = new_moment(name, state["exp_avg"], new)
            state["exp_avg_sq"] = new_moment(name, state["exp_avg_sq"], new)
        if state is not None:
            optimizer.state[new] = state
        group["params"][0] = new
        out[name] = new
    return out


class CurveTopology:
    """Mixin for GaussianCurveModel: the reference's training-time surgery on the curve set."""

    _GROUPS = (("f_dc", "_features_dc"), ("f_rest", "_features_rest"), ("opacity", "_opacity"), ("width", "_width"),
               ("curve_points", "_curve_points"), ("mask", "_mask"))

    # ---- optimizer -----------------------------------------------------------------------------------
    def training_setup(self, training_args):
        """Adam over the six per-curve tensors with the reference's group names and rates
        (gaussian_curve_model.py:200-232); the curve-point rate follows `update_learning_rate`."""
        dev = self._curve_points.device
        P = self._curve_points.shape[0] * self.n_gaussians
        self.denom = torch.zeros((P, 1), device=dev)
        self.xyz_gradient_accum = torch.zeros((P, 1), device=dev)
        g = lambda k, d=0.0: getattr(training_args, k, d)
        rates = {"f_dc": g("feature_lr", 0.0025), "f_rest": g("feature_lr", 0.0025) / 20.0, "opacity": g("opacity_lr", 0.025),
                 "width": g("scaling_lr", 0.005), "curve_points": g("lr_curve_points_init", 0.0005),
                 "mask": g("mask_lr", 0.01)}
        groups = [{"params": [getattr(self, attr)], "lr": rates[name], "name": name} for name, attr in self._GROUPS]
        self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        self.curve_scheduler_args = get_expon_lr_func(
            lr_init=g("lr_curve_points_init", 0.0005), lr_final=g("lr_curve_points_final", 0.000005),
            lr_delay_mult=g("position_lr_delay_mult", 0.01), max_steps=g("position_lr_max_steps", 30000))
        # per-image exposure (train.py:228-229 steps this optimizer every iteration; unused unless train_test_exp)
        if not isinstance(getattr(self, "_exposure", None), nn.Parameter):
            self._exposure = nn.Parameter(torch.eye(3, 4, device=dev)[None].clone().requires_grad_(True))
        self.exposure_optimizer = torch.optim.Adam([self._exposure])
        self.exposure_scheduler_args = get_expon_lr_func(
            g("exposure_lr_init", 0.01), g("exposure_lr_final", 0.001), lr_delay_steps=g("exposure_lr_delay_steps", 0),
            lr_delay_mult=g("exposure_lr_delay_mult", 0.0), max_steps=g("iterations", 10000))
        return self.optimizer

    def update_learning_rate(self, iteration):
        if getattr(self, "pretrained_exposures", None) is None and getattr(self, "exposure_optimizer", None) is not None:
            for group in self.exposure_optimizer.param_groups:
                group["lr"] = self.exposure_scheduler_args(iteration)
        for group in self.optimizer.param_groups:
            if group["name"] == "curve_points":
                group["lr"] = self.curve_scheduler_args(iteration)
                return group["lr"]

    def _adopt(self, tensors: Dict[str, nn.Parameter]) -> None:
        # every replacement of a parameter tensor bumps the version: anything that captured the old tensors
        # (a CUDA graph of the step, trainer.TrainLoop) knows it has to be rebuilt
        self.topology_version = getattr(self, "topology_version", 0) + 1
        for name, attr in self._GROUPS:
            if name in tensors:
                setattr(self, attr, tensors[name])

    def _prune_optimizer(self, keep):
        return _edit_groups(self.optimizer, None, lambda _, p: p[keep], lambda _, m, __: m[keep])

    def cat_tensors_to_optimizer(self, tensors_dict):
        return _edit_groups(self.optimizer, None, lambda k, p: torch.cat((p, tensors_dict[k]), dim=0),
                            lambda k, m, _: torch.cat((m, torch.zeros_like(tensors_dict[k])), dim=0))

    def replace_tensor_to_optimizer(self, tensor, name):
        return _edit_groups(self.optimizer, (name,), lambda _, __: tensor, lambda _, __, new: torch.zeros_like(new))

    def add_densification_stats(self, viewspace_point_tensor, update_filter):
        """scene/gaussian_model.py:618-620 (update_filter: boolean (P,) or index tensor)."""
        g = viewspace_point_tensor.grad
        if update_filter.dtype == torch.bool:
            # masked adds: same sums as the reference's boolean indexing without its host-synchronising nonzero()
            f = update_filter.view(-1, 1).to(self.denom.dtype)
            self.xyz_gradient_accum += torch.norm(g[:, :2], dim=-1, keepdim=True) * f
            self.denom += f
            return
        self.xyz_gradient_accum[update_filter] += torch.norm(g[update_filter, :2], dim=-1, keepdim=True)
        self.denom[update_filter] += 1

    # ---- opacity resets ---------------------------------------------------------------------------------
    def reset_opacity(self):
        new = self.inverse_opacity_activation(torch.clamp_max(self.get_curve_opacity.detach(), 0.1))
        self._adopt(self.replace_tensor_to_optimizer(new, "opacity"))

    def fix_opacity(self):
        new = self.inverse_opacity_activation(torch.clamp_min(self.get_curve_opacity.detach(), 0.6))
        self._adopt(self.replace_tensor_to_optimizer(new, "opacity"))
        self._opacity.requires_grad = False
        for group in self.optimizer.param_groups:
            if group["name"] == "opacity":
                group["lr"] = 0.0

    # ---- prune / append -----------------------------------------------------------------------------------
    def prune_curves(self, mask):
        """Drop the curves where mask is True, with their Adam moments and per-Gaussian statistics
        (gaussian_curve_model.py:283-305)."""
        keep = ~mask
        self._adopt(self._prune_optimizer(keep))
        keep_pts = keep[:, None].expand(-1, self.n_gaussians).reshape(-1)
        self.xyz_gradient_accum = self.xyz_gradient_accum[keep_pts]
        self.denom = self.denom[keep_pts]
        self.is_bezier = self.is_bezier[keep]
        self.max_radii2D = self.max_radii2D[keep_pts]
        tmp = getattr(self, "tmp_radii", None)
        if tmp is not None and tmp.shape[0] == keep_pts.shape[0]:
            self.tmp_radii = tmp[keep_pts]
        self.prepare_scaling_rot()

    def densification_postfix(self, new_curve_points, new_features_dc, new_features_rest, new_opacities, new_widths,
                              new_masks, new_is_bezier):
        """Append curves; the per-Gaussian statistics restart from zero (gaussian_curve_model.py:307-326)."""
        self._adopt(self.cat_tensors_to_optimizer({
            "curve_points": new_curve_points, "f_dc": new_features_dc, "f_rest": new_features_rest,
            "opacity": new_opacities, "width": new_widths, "mask": new_masks}))
        self.is_bezier = torch.cat((self.is_bezier, new_is_bezier))
        dev = self._curve_points.device
        P = self._curve_points.shape[0] * self.n_gaussians
        self.xyz_gradient_accum = torch.zeros((P, 1), device=dev)
        self.denom = torch.zeros((P, 1), device=dev)
        self.max_radii2D = torch.zeros(P, device=dev)

    def densify_and_split_curve(self, selected, t, N=2):
        """Replace every selected curve by its two halves at parameter t ((K,1), K = selected.sum());
        the halves inherit width, opacity and mask (gaussian_curve_model.py:330-347)."""
        assert N == 2
        k = int(selected.sum())
        rep = lambda x: x.detach()[selected].repeat(N, *([1] * (x.dim() - 1)))
        left, right = de_casteljau_split(self._curve_points.detach()[selected], t, self.is_bezier[selected])
        self.densification_postfix(torch.cat((left, right), dim=0), rep(self._features_dc), rep(self._features_rest),
                                   rep(self._opacity), rep(self._width), rep(self._mask), self.is_bezier[selected].repeat(N))
        gone = torch.cat((selected, torch.zeros(N * k, device=selected.device, dtype=torch.bool)))
        self.prune_curves(gone)

    def densify_and_prune(self, max_grad, min_opacity, extent, max_screen_size, radii):
        """Split each curve whose largest mean screen-space gradient reaches max_grad at the sample where it
        occurs, then drop curves fainter than min_opacity (gaussian_curve_model.py:349-365)."""
        grads = self.xyz_gradient_accum / self.denom
        grads[grads.isnan()] = 0.0
        self.tmp_radii = radii
        per_curve = torch.norm(grads.view(-1, self.n_gaussians, grads.shape[-1]), dim=-1)
        top, where = per_curve.max(dim=1)
        selected = top >= max_grad
        if bool(selected.any()):
            self.densify_and_split_curve(selected, self.sample_t[where[selected]].squeeze(-1))
        self.prune_curves((self.get_curve_opacity < min_opacity).reshape(-1))

    def curve_split_curvature(self, threshold_angle=20, threshold_radian_skip=30):
        """Split at the sharpest bend the curves whose main axis turns by more than threshold_angle degrees
        between neighbouring samples (or threshold_radian_skip between samples two apart)
        (gaussian_curve_model.py:372-387)."""
        n = self.n_gaussians
        axis = self.get_rotation_matrix[..., 0].detach().view(-1, n, 3)
        ang = torch.acos((axis[:, :-1] * axis[:, 1:]).sum(-1).clamp(-1, 1))
        ang2 = torch.acos((axis[:, :-2] * axis[:, 2:]).sum(-1).clamp(-1, 1))
        split = (ang.max(dim=-1).values > math.radians(threshold_angle)) | \
                (ang2.max(dim=-1).values > math.radians(threshold_radian_skip))
        where = ang.argmax(dim=-1)
        end_t = self.sample_t[where] + 0.5 / n
        self.densify_and_split_curve(split, end_t[split].squeeze(-1))
        self.prepare_scaling_rot()

    def only_prune(self, min_opacity, mask_threshold):
        """Drop curves that are fully masked out, too faint, or shorter than 1e-2 in summed sample scale
        (gaussian_curve_model.py:427-435)."""
        masked = (torch.sigmoid(self._mask.detach()) <= mask_threshold).all(dim=1).reshape(-1)
        faint = (self.get_curve_opacity < min_opacity).reshape(-1)
        short = self._scaling[:, 0].detach().reshape(-1, self.n_gaussians).sum(-1) < 1e-2
        self.prune_curves(masked | faint | short)

    def mask_trim_split(self, mask_threshold):
        """Cut the masked-out samples off both ends of every curve and stretch the surviving part of its mask
        logits back over all n samples (gaussian_curve_model.py:437-459)."""
        n = self.n_gaussians
        m = self._mask.detach().view(-1, n)
        valid = torch.sigmoid(m) > mask_threshold
        start = valid.int().argmax(dim=1)
        end = n - 1 - valid.flip(1).int().argmax(dim=1)
        t = self.sample_t.view(-1)
        from_t = (t[start] - 0.5 / n)[:, None]
        end_t = (t[end] + 0.5 / n)[:, None]
        trimmed = de_casteljau_trim(self._curve_points.detach(), from_t, end_t, self.is_bezier)
        touched = (start != 0) | (end != n - 1)
        new_mask = torch.where(touched[:, None], resample_mask_rows(m, start, torch.maximum(end, start)), m)
        self._adopt(self.replace_tensor_to_optimizer(new_mask.view(-1, n, 1).contiguous(), "mask"))
        self._adopt(self.replace_tensor_to_optimizer(trimmed.contiguous(), "curve_points"))
        self.prepare_scaling_rot()

    def fit_curve_to_line(self, threshold=0.002, threshold_max=0.004, sample_num=100):
        """Re-label as straight segments the Beziers whose sample_num points lie within `threshold` (mean) and
        `threshold_max` (max) of their least-squares line (gaussian_curve_model.py:590-626 with
        edge_extraction/fitting.py:74-97), for all curves at once: batched 3x3 eigen-decomposition instead of a
        per-curve numpy loop. As in the reference the control points keep their values (its write of the fitted
        end points goes to a temporary copy, :612-613), `is_bezier` flips and the Adam moments of the curve
        points restart. Returns the number of curves re-labelled."""
        cp = self._curve_points.detach()
        t = torch.linspace(0, 1, sample_num, device=cp.device)[None, :, None]
        p0, p1, p2, p3 = (cp[:, i][:, None, :] for i in range(4))
        pts = (1 - t) ** 3 * p0 + 3 * (1 - t) ** 2 * t * p1 + 3 * (1 - t) * t ** 2 * p2 + t ** 3 * p3      # (B,S,3)
        mean = pts.mean(dim=1, keepdim=True)
        c = pts - mean
        cov = c.transpose(1, 2) @ c / sample_num
        direction = torch.linalg.eigh(cov).eigenvectors[..., -1]                 # largest eigenvalue
        direction = direction / direction.norm(dim=-1, keepdim=True)
        along = (c * direction[:, None, :]).sum(-1, keepdim=True)               # projections lie in [t_min, t_max]
        dist = (c - along * direction[:, None, :]).norm(dim=-1)
        straight = (dist.mean(dim=1) < threshold) & (dist.max(dim=1).values < threshold_max) & self.is_bezier
        k = int(straight.sum())
        if k:
            self.is_bezier = self.is_bezier & ~straight
            self._adopt(self.replace_tensor_to_optimizer(cp.clone(), "curve_points"))
        return k

    # ---- merging -----------------------------------------------------------------------------------------------
    def sample_curve_points(self, sample_num=100):
        """(B, sample_num, 3) points at t = linspace(0, 1): cubic Beziers, or the chord for straight segments
        (get_curve_gaussians, gaussian_curve_model.py:70-78)."""
        cp = self._curve_points.detach()
        t = torch.linspace(0, 1, sample_num, device=cp.device)[None, :, None]
        p0, p1, p2, p3 = (cp[:, i][:, None, :] for i in range(4))
        bez = (1 - t) ** 3 * p0 + 3 * (1 - t) ** 2 * t * p1 + 3 * (1 - t) * t ** 2 * p2 + t ** 3 * p3
        return torch.where(self.is_bezier[:, None, None], bez, (1 - t) * p0 + t * p3)

    def _bezier_merge_candidates(self, distance_threshold, similarity_threshold, rows=2048):
        """Edges (i, j, confidence) of the reference's end-point adjacency (gaussian_curve_model.py:469-484): some
        end of i within 2*distance_threshold of some end of j with |cos| of the end tangents above the threshold;
        confidence = the largest |cos| over the four end pairings. Built in row blocks, so the (2B)^2 matrices of
        the reference never exist."""
        cp = self._curve_points.detach()
        B = cp.shape[0]
        ends = (cp[:, 0], cp[:, 3])
        tang = [cp[:, 1] - cp[:, 0], cp[:, 2] - cp[:, 3]]
        tang = [v / (v.norm(dim=-1, keepdim=True) + 1e-6) for v in tang]
        out = []
        for r0 in range(0, B, rows):
            r1 = min(B, r0 + rows)
            adj = torch.zeros((r1 - r0, B), dtype=torch.bool, device=cp.device)
            conf = torch.zeros((r1 - r0, B), dtype=cp.dtype, device=cp.device)
            for a in range(2):
                for b in range(2):
                    sim = (tang[a][r0:r1] @ tang[b].T).abs()
                    adj |= (torch.cdist(ends[a][r0:r1], ends[b]) < 2 * distance_threshold) & (sim > similarity_threshold)
                    conf = torch.maximum(conf, sim)
            i, j = adj.nonzero(as_tuple=True)
            out.append(torch.stack([(i + r0).double(), j.double(), conf[i, j].double()], dim=1))
        return torch.cat(out).cpu() if out else torch.zeros((0, 3), dtype=torch.float64)

    def merge_curves(self, distance_threshold=0.02, similarity_threshold=0.97, sample_num=100, ransac_thresh=0.005):
        """Join Beziers that continue each other end to end, and straight segments that are collinear and touching
        (gaussian_curve_model.py:459-588).

        Beziers: greedy pairing in index order - every curve takes its most tangent-aligned unpaired neighbour -
        then one cubic is fitted through the 2*sample_num points of a pair, ordered along their principal axis,
        and replaces the pair if its RMS residual stays below distance_threshold. Segments: connected components of
        (distance <= distance_threshold, |cos| >= similarity_threshold) become the extent of their points along
        the principal axis. A merged curve gets the mean width / opacity logit of its parts and a fresh mask.
        Differences from the reference: its RANSAC consensus step (`ransac_thresh`, unseeded) is replaced by using
        every sampled point; the least-squares cubic is solved in closed form (the parametrisation is fixed, so the
        fit is linear) instead of scipy's iterative curve_fit; pairs and components are processed batched.
        Returns the number of curves removed."""
        cp = self._curve_points.detach()
        dev, B = cp.device, cp.shape[0]
        if B == 0:
            return 0
        pts = self.sample_curve_points(sample_num)
        is_bez = self.is_bezier.cpu().tolist()
        merge_mask = torch.zeros(B, dtype=torch.bool, device=dev)
        new_cp, new_isb, parts = [], [], []

        # -- Beziers: greedy pairs (host walk over the sparse candidate list, as the reference walks its matrix)
        edges = self._bezier_merge_candidates(distance_threshold, similarity_threshold).tolist()
        nbrs: Dict[int, list] = {}
        for i, j, c in edges:
            nbrs.setdefault(int(i), []).append((int(j), c))
        taken, pairs = set(), []
        for i in range(B):
            if i in taken or not is_bez[i]:
                continue
            cand = [(j, c) for j, c in nbrs.get(i, []) if j not in taken and j != i and is_bez[j]]
            if not cand:
                continue
            best = max(cand, key=lambda jc: jc[1])[0]       # first maximum in index order, like max() over the list
            taken.update((i, best))
            pairs.append((i, best))
        if pairs:
            pr = torch.tensor(pairs, device=dev)
            P = torch.cat([pts[pr[:, 0]], pts[pr[:, 1]]], dim=1).double()                  # (K, 2S, 3)
            C = P - P.mean(dim=1, keepdim=True)
            axis = torch.linalg.eigh(C.transpose(1, 2) @ C).eigenvectors[..., -1]        # principal direction
            order = (C * axis[:, None, :]).sum(-1).argsort(dim=1)
            P = torch.gather(P, 1, order[:, :, None].expand(-1, -1, 3))
            n = P.shape[1]
            t = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
            basis = torch.stack([(1 - t) ** 3, 3 * (1 - t) ** 2 * t, 3 * (1 - t) * t ** 2, t ** 3], dim=1)   # (n, 4)
            ctrl = torch.linalg.pinv(basis) @ P                                          # least squares, (K, 4, 3)
            rmse = ((P - basis @ ctrl) ** 2).sum(-1).mean(dim=1).sqrt()
            ok = rmse <= distance_threshold
            for k in ok.nonzero().view(-1).tolist():
                merge_mask[list(pairs[k])] = True
                new_cp.append(ctrl[k].to(cp.dtype))
                new_isb.append(True)
                parts.append(list(pairs[k]))

        # -- straight segments: connected components of "touching and parallel"
        line_idx = (~self.is_bezier).nonzero().view(-1)
        L = int(line_idx.numel())
        if L > 1:
            a, b = cp[line_idx, 0].double(), cp[line_idx, 3].double()
            d = b - a

            def seg_to_points(q):            # [i, j] = distance from segment i to point q_j
                u = (((q[None, :, :] - a[:, None, :]) * d[:, None, :]).sum(-1) / (d * d).sum(-1)[:, None]).clamp(0, 1)
                return (a[:, None, :] + u[..., None] * d[:, None, :] - q[None, :, :]).norm(dim=-1)

            upper = torch.triu(torch.minimum(seg_to_points(a), seg_to_points(b)), diagonal=1)
            dist = upper + upper.T           # the reference fills i < j and mirrors (edge_extraction/merging.py:84-107)
            dn = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-300)
            adj = (dist <= distance_threshold) & ((dn @ dn.T).abs() >= similarity_threshold)
            adj |= torch.eye(L, dtype=torch.bool, device=dev)
            labels = torch.arange(L, device=dev)
            while True:                      # min-label propagation: a component is named after its first member
                nxt = torch.where(adj, labels[None, :], L).min(dim=1).values
                nxt = torch.minimum(nxt, labels)
                if torch.equal(nxt, labels):
                    break
                labels = nxt
            for lab in labels.unique().tolist():
                members = (labels == lab).nonzero().view(-1)
                if members.numel() < 2:
                    continue
                which = line_idx[members]
                merge_mask[which] = True
                Q = pts[which].reshape(-1, 3).double()
                mean = Q.mean(dim=0)
                Cq = Q - mean
                axis = torch.linalg.eigh(Cq.T @ Cq / Q.shape[0]).eigenvectors[:, -1]
                proj = Cq @ axis
                seg = torch.zeros((4, 3), dtype=cp.dtype, device=dev)    # inner control points of a segment are unused
                seg[0], seg[3] = (mean + proj.min() * axis).to(cp.dtype), (mean + proj.max() * axis).to(cp.dtype)
                new_cp.append(seg)
                new_isb.append(False)
                parts.append(which.tolist())

        removed = int(merge_mask.sum())
        if removed:
            opac = torch.stack([self._opacity.detach()[p].mean(dim=0) for p in parts])
            width = torch.stack([self._width.detach()[p].mean(dim=0) for p in parts])
            k = len(parts)
            f_dc = self._features_dc.detach()[0:1].repeat(k, 1, 1, 1)
            f_rest = self._features_rest.detach()[0:1].repeat(k, 1, 1, 1)
            masks = torch.ones((k,) + tuple(self._mask.shape[1:]), dtype=self._mask.dtype, device=dev)
            self.prune_curves(merge_mask)
            self.densification_postfix(torch.stack(new_cp), f_dc, f_rest, opac, width, masks,
                                       torch.tensor(new_isb, dtype=torch.bool, device=dev))
            self.prepare_scaling_rot()
        return removed
