"""Host-side keep-rate schedule of the fine-tune loop (reference audiomae/engine_finetune.py:29-53).

During fine-tuning the reference does not prune at a fixed rate from the first step: until ``shrink_start_epoch`` every
block keeps all tokens, then the keep rate of the ``drop_loc`` blocks follows half a cosine from ``max_keep_rate`` down
to ``base_keep_rate`` over the shrink epochs, and afterwards the model's own defaults apply (``None``).  The returned
tuple is what ``model(x, keep_rate_list=...)`` takes; every distinct value costs one CUDA-graph capture when
``use_cuda_graph`` is on (the engine keeps the eight most recent schedules).
"""
import math
from typing import Optional, Sequence, Tuple


def get_scheduled_keep_rate_list(iters: int, epoch: int, shrink_start_epoch: int, total_epochs: int, ITERS_PER_EPOCH: int,
                                 base_keep_rate: float = 0.5, max_keep_rate: float = 1, num_blocks: int = 12,
                                 drop_loc: Sequence[int] = (3, 6, 9)) -> Optional[Tuple[float, ...]]:
    if epoch < shrink_start_epoch:
        return (1.0,) * num_blocks                        # do not drop any token yet
    if epoch >= total_epochs:
        return None                                        # the model follows its default keep rates
    total_iters = ITERS_PER_EPOCH * (total_epochs - shrink_start_epoch)
    iters = iters - ITERS_PER_EPOCH * shrink_start_epoch
    target = base_keep_rate + (max_keep_rate - base_keep_rate) * (math.cos(iters / total_iters * math.pi) + 1) * 0.5
    return tuple(target if i in drop_loc else 1.0 for i in range(num_blocks))
