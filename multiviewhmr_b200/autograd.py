"""Backward of unproject+aggregate w.r.t. the feature maps (SURVEY.md §8(f)
rank 1).  Not built yet: training through the fused op fails loudly instead of
silently detaching."""


def unprojection_with_grad(features, proj_matricies, coord_volumes, aggregation_method):
    raise NotImplementedError(
        "multiviewhmr_b200: backward of the fused unprojection is not implemented yet; "
        "call under torch.no_grad() or detach the feature maps")
