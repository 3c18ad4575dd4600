"""autograd bridge: the whole generator is ONE ``torch.autograd.Function`` whose backward is the
hand-scheduled kernel sequence of ``engine_train.TrainEngine``.  Parameters enter as ordinary
inputs, so their gradients come back through autograd's AccumulateGrad nodes -- optimizers, DDP
hooks and ``.grad`` semantics are exactly those of the reference's nn.Conv2d parameters."""
from __future__ import annotations

import torch


class GeneratorFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gen, x, *params):  # noqa: ARG004 (params only establish graph edges)
        eng = gen._get_engine(train=True)
        out, bufs = eng.forward_train(x.detach())
        ctx.gen, ctx.bufs, ctx.generation = gen, bufs, eng.generation
        ctx.save_for_backward(x)
        ctx.mark_non_differentiable()
        return out

    @staticmethod
    def backward(ctx, gout):
        (x,) = ctx.saved_tensors
        eng = ctx.gen._get_engine(train=True)
        need_x = ctx.needs_input_grad[1]
        gx, grads = eng.backward(ctx.bufs, ctx.generation, x, gout, need_x)
        out = [None, gx]
        for need, g in zip(ctx.needs_input_grad[2:], grads):
            out.append(g if need else None)
        return tuple(out)


def generator_apply(gen, x: torch.Tensor) -> torch.Tensor:
    return GeneratorFunction.apply(gen, x, *gen.parameters())
