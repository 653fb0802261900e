"""Drop-in for the live part of the reference's lib/hough_voting.py (``HoughVotingLayer.forward``, :41-63)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .ransac_voting_gpu_layer import ransac_voting_gpu as rvg


class HoughVotingLayer(nn.Module):

    def __init__(self, HPARAM):
        super().__init__()
        self.HPARAM = HPARAM

    def forward(self, agg_data, *, idxs=None, select_mask=None):
        """``agg_data['xy']`` [N,2,h,w] masked direction field + ``agg_data['instance_masks']`` [N,h,w]
        -> adds 'hypothesis', 'pruned_hypothesis' [N,1,2], replaces 'xy' by the centre [N,2] (col,row) and
        keeps the dense field as 'xy_mask'.  Keyword-only ``idxs`` [N,hn,1,2] int32 fixes the sampled pairs."""
        uv_img = agg_data["xy"]
        mask = agg_data["instance_masks"]
        reshaped_uv_img = torch.unsqueeze(uv_img.permute(0, 2, 3, 1), dim=3)      # [N,h,w,1,2] view (hough_voting.py:51)
        output = rvg.ransac_voting_layer_v3(mask=mask, vertex=reshaped_uv_img,
                                            round_hyp_num=self.HPARAM.HV_NUM_OF_HYPOTHESES,
                                            idxs=idxs, select_mask=select_mask)
        good_output = torch.squeeze(output, dim=1)
        agg_data.update({"hypothesis": output, "pruned_hypothesis": output, "xy": good_output, "xy_mask": uv_img})
        return agg_data
