"""Drop-in for reference ``utils/rotation_conversions.py`` (a PyTorch3D excerpt): Euler angles -> rotation matrices.

Device-agnostic host-side helper with the reference's signatures and error behaviour (ValueError on malformed
input, rotation_conversions.py:51-59).  Inside the training step the same rotation R = Rx(a) Ry(b) Rx(g) is built
in registers by the fused geometry kernels (csrc/geom.cuh: make_rotation); this module serves callers outside it.
"""
import torch

_AXES = {"X": 0, "Y": 1, "Z": 2}


def _axis_angle_rotation(axis: str, angle):
    """Rotation about one coordinate axis for every angle in `angle` -> (..., 3, 3)."""
    if axis not in _AXES:
        raise ValueError(f"Invalid letter {axis} in convention string.")
    c, s = torch.cos(angle), torch.sin(angle)
    R = torch.zeros(angle.shape + (3, 3), dtype=angle.dtype, device=angle.device)
    k = _AXES[axis]
    i, j = (k + 1) % 3, (k + 2) % 3          # the plane being rotated, in cyclic order
    R[..., k, k] = 1.0
    R[..., i, i] = c
    R[..., j, j] = c
    R[..., i, j] = -s
    R[..., j, i] = s
    return R


def euler_angles_to_matrix(euler_angles, convention: str):
    """euler_angles (..., 3) in radians, convention e.g. 'XYZ' -> (..., 3, 3) = R_c0 @ R_c1 @ R_c2."""
    if euler_angles.dim() == 0 or euler_angles.shape[-1] != 3:
        raise ValueError("Invalid input euler angles.")
    if len(convention) != 3:
        raise ValueError("Convention must have 3 letters.")
    if convention[1] in (convention[0], convention[2]):
        raise ValueError(f"Invalid convention {convention}.")
    for letter in convention:
        if letter not in _AXES:
            raise ValueError(f"Invalid letter {letter} in convention string.")
    out = None
    for k, letter in enumerate(convention):
        R = _axis_angle_rotation(letter, euler_angles[..., k])
        out = R if out is None else torch.matmul(out, R)
    return out
