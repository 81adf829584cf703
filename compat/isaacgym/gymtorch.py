"""isaacgym.gymtorch: engine buffers already ARE torch views, so wrapping / unwrapping is the identity."""


def wrap_tensor(t):
    return t


def unwrap_tensor(t):
    return t
