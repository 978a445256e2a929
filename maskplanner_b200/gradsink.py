"""Direct gradient placement for the data-parallel step.

The reference is single-process (train_maskplanner.py:220-221: ``loss.backward(); opt.step()``).  Under data parallelism
the flat gradient buffer is what the all-reduce sends, and round 1 filled it the autograd way: a memset of
the 105 MB buffer, ~80 ``AccumulateGrad`` read-modify-write kernels into its views, then two scaling passes.  Every
parameter of the model receives its gradient from exactly one kernel of this library per step (the weight-gradient GEMMs,
the BatchNorm finalize kernels, the head kernels), so those kernels can write into the flat buffer themselves:

    register(param, view)   the Trainer maps each parameter to its 16-byte aligned view of the flat buffer
    lookup(param)           autograd Functions ask for the view inside backward(); when one exists the kernel writes there
                            and the Function returns None for that parameter (autograd then leaves ``.grad`` -- the same
                            view -- alone: overwrite semantics, one gradient per step)

Nothing is registered on a single GPU: autograd then receives ordinary gradient tensors.
"""
_SINKS = {}
HOOKS = {"heads_done": None}      # called at the end of the heads' backward (the Trainer launches the head-bucket collective)


def register(param, view):
    _SINKS[param.data_ptr()] = view


def clear():
    _SINKS.clear()
    HOOKS["heads_done"] = None


def lookup(param):
    if not _SINKS or param is None:
        return None
    return _SINKS.get(param.data_ptr())


def active():
    return bool(_SINKS)
