"""Training path (forward with saved activations + hand-written backward).  Not built yet."""


def dpot_forward_train(net, x):
    raise NotImplementedError(
        "dpot_b200: the backward kernels are not built yet; run DPOTNet under torch.no_grad() / model.eval() "
        "with requires_grad_(False) inputs.  (No PyTorch fallback is provided on purpose.)")
