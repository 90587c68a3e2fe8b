"""Reads the reference's process-global `flags` object (phc/utils/flags.py) when the reference is
importable; the step path only needs `im_eval` (humanoid_im.py:1186: use the mean-distance reset
criterion when `flags.im_eval and not self.strict_eval`)."""


def reference_flags():
    try:
        from phc.utils.flags import flags  # type: ignore
        return flags
    except Exception:
        return None


def im_eval_mean_reset(task) -> bool:
    fl = reference_flags()
    if fl is None:
        return bool(getattr(task, "_pulse_im_eval", False)) and not bool(getattr(task, "strict_eval", False))
    return bool(getattr(fl, "im_eval", False)) and not bool(getattr(task, "strict_eval", False))


def flags_test() -> bool:
    """`flags.test` (evaluation run: the distillation teacher is skipped, humanoid_im_distill.py:151; the VAE uses z = mu,
    amp_network_z_builder.py:94-95)."""
    fl = reference_flags()
    return bool(getattr(fl, "test", False)) if fl is not None else False
