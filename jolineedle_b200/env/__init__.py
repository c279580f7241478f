from .common import Action, get_actions_info  # noqa: F401
