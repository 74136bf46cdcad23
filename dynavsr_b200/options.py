"""Option-dict helpers the model wrappers rely on (codes/options/options.py:108-123): a dict whose missing keys read as
``None``.  YAML parsing / path derivation stay with the reference's ``options.parse`` -- its output feeds ``create_model``
unchanged."""


class NoneDict(dict):
    def __missing__(self, key):
        return None


def dict_to_nonedict(opt):
    if isinstance(opt, dict):
        return NoneDict(**{k: dict_to_nonedict(v) for k, v in opt.items()})
    if isinstance(opt, list):
        return [dict_to_nonedict(v) for v in opt]
    return opt
