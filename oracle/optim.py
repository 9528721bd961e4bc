"""ORACLE — test infrastructure only (never imported by the product path).

CPU restatement (numpy float32, one explicit elementwise expression per reference line) of

  RMSprop.step                       /root/reference/utils/rmsprop.py:67-129
  ExponentialMovingAverage.forward   /root/reference/utils/optim.py:53-64
  ExponentialMovingAverage.adjust_momentum   /root/reference/utils/optim.py:118-128
  cal_l2_loss('mnas') gradient       /root/reference/utils/optim.py:177-200
  _allreduce_coalesced (mean)        /root/reference/utils/distributed.py:131-139

Pinning: the reference has no test for RMSprop ("parity unpinned" upstream); the
EMA rule is pinned by the reference's own known-answer vectors
(/root/reference/tests/utils/optim_test.py:113-128,161-172: decay 0.25 and the num_updates rule
min(decay,(1+t)/(10+t)) -> 0.181818 at t=1) reproduced in tests/test_oracle_optim.py, and both are
pinned against the live reference classes by oracle/make_golden.py -> tests/golden/optim.pt.
"""
import numpy as np

f32 = np.float32


def rmsprop_step(p, g, sq, mom, lr, alpha=0.99, eps=1e-8, eps_inside_sqrt=False, momentum=0.0,
                 weight_decay=0.0, grad_avg=None):
    """One RMSprop.step() on flat float32 arrays; returns (p, sq, mom, grad_avg) (new arrays).

    Follows utils/rmsprop.py:97-127 line by line, in float32 like the reference tensors.
    """
    p, g, sq = p.astype(f32), g.astype(f32), sq.astype(f32)
    if weight_decay != 0:                                     # :99-100
        g = g + f32(weight_decay) * p
    sq = sq * f32(alpha) + f32(1 - alpha) * g * g             # :102
    if grad_avg is not None:                                  # centered, :104-112
        grad_avg = grad_avg.astype(f32) * f32(alpha) + f32(1 - alpha) * g
        if eps_inside_sqrt:
            avg = np.sqrt(sq - grad_avg * grad_avg + f32(eps))
        else:
            avg = np.sqrt(sq - grad_avg * grad_avg) + f32(eps)
    else:
        if eps_inside_sqrt:                                   # :114-115
            avg = np.sqrt(sq + f32(eps))
        else:                                                 # :116-117
            avg = np.sqrt(sq) + f32(eps)
    if momentum > 0:                                          # :119-122
        mom = mom.astype(f32) * f32(momentum) + g / avg
        p = p - f32(lr) * mom
    else:                                                     # :123-124
        p = p - f32(lr) * (g / avg)
    return p.astype(f32), sq.astype(f32), (mom.astype(f32) if mom is not None else None), grad_avg


def ema_momentum(decay, num_updates=None):
    """utils/optim.py:56-60."""
    if num_updates is None:
        return decay
    return min(decay, (1.0 + num_updates) / (10.0 + num_updates))


def ema_update(shadow, x, decay, num_updates=None):
    """shadow <- m*shadow + (1-m)*x   (utils/optim.py:63-64)."""
    m = ema_momentum(decay, num_updates)
    return (shadow.astype(f32) * f32(m) + f32(1.0 - m) * x.astype(f32)).astype(f32)


def adjust_momentum(momentum, steps_multi):
    """utils/optim.py:128."""
    return momentum ** (1.0 / steps_multi)


def l2_decay_mask(named_shapes):
    """Which parameters cal_l2_loss(method='mnas') regularises (utils/optim.py:180-191):
    every 4-D / 2-D weight and the classifier bias; BN gamma/beta are excluded."""
    mask = {}
    for name, shape in named_shapes:
        if len(shape) in (4, 2):
            mask[name] = True
        else:
            assert len(shape) == 1
            mask[name] = "classifier" in name
    return mask


def l2_grad(p, weight_decay):
    """d/dp [0.5 * wd * sum(p^2)] = wd * p   (utils/optim.py:193-200)."""
    return f32(weight_decay) * p.astype(f32)


def allreduce_mean(grads_per_rank):
    """SUM over ranks then divide by world size (utils/distributed.py:135-136)."""
    tot = grads_per_rank[0].astype(f32).copy()
    for g in grads_per_rank[1:]:
        tot = tot + g.astype(f32)
    return tot / f32(len(grads_per_rank))
