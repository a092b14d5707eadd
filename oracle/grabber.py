"""ORACLE (test infrastructure; never imported by the product): MouseGrabber restated on the CPU.

Reference: /root/reference/Velvet/MouseGrabber.hpp
  * FindClosestVertexToRay, L92-110: host loop over every position; distanceToView = dot(dir, p - origin),
    distanceToRay = length(cross(dir, p - origin)); the particle with the smallest distanceToView among those with
    distanceToRay < particleDiameter, first index on ties (strict `<`, minDistanceToView starts at FLT_MAX);
  * HandleMouseInteraction, L40-63: on a hit the particle's inverse mass is saved and set to 0; release restores it;
  * UpdateGrappedVertex, L66-79: mousePos = origin + dir * distanceToOrigin; target = Lerp(mousePos, curPos, 0.8)
    (Helper.hpp L36-40: a * value2 + (1 - a) * value1); positions[id] = target; velocities[id] = (target - curPos) / fixedDeltaTime.
All arithmetic in fp32, one rounding per operation, in glm's evaluation order (glm::dot sums left to right, glm::cross is
a.y*b.z - b.y*a.z, ...).  The camera unprojection that produces the ray (L112-130) is outside the path.
Parity pinned by: known-answer cases in tests/test_oracle_cpu.py (a ray through a chosen vertex, occlusion order, a miss).
"""
import numpy as np

F = np.float32
FLT_MAX = np.finfo(np.float32).max
FIXED_DELTA_TIME = F(1.0 / 60.0)  # Timer::fixedDeltaTime(), Timer.hpp


def find_closest_vertex_to_ray(positions, origin, direction, particle_diameter):
    """(index or -1, distanceToView of the pick or FLT_MAX).  Vectorised: every product / sum below is one fp32 operation on
    the whole array, which rounds exactly like the reference's scalar loop."""
    p = np.ascontiguousarray(positions, F).reshape(-1, 3)
    o, d = np.asarray(origin, F), np.asarray(direction, F)
    rel = p - o
    with np.errstate(all="ignore"):
        view = (d[0] * rel[:, 0] + d[1] * rel[:, 1]) + d[2] * rel[:, 2]
        cx = d[1] * rel[:, 2] - rel[:, 1] * d[2]
        cy = d[2] * rel[:, 0] - rel[:, 2] * d[0]
        cz = d[0] * rel[:, 1] - rel[:, 0] * d[1]
        to_ray = np.sqrt((cx * cx + cy * cy) + cz * cz)
        ok = (to_ray < F(particle_diameter)) & (view < FLT_MAX)
    if not ok.any():
        return -1, FLT_MAX
    masked = np.where(ok, view, np.inf)
    i = int(np.argmin(masked))  # first index of the minimum, like the loop's strict `<`
    return i, F(view[i])


class MouseGrabber:
    """Operates on numpy views of a solver's positions [n,3], velocities [n,3] and invMasses [n] (the oracle's own arrays)."""

    def __init__(self, positions, velocities, inv_masses, particle_diameter):
        self.positions = positions.reshape(-1, 3)
        self.velocities = velocities.reshape(-1, 3)
        self.inv_masses = inv_masses
        self.particle_diameter = F(particle_diameter)
        self.grabbing = False
        self.index = -1
        self.distance = FLT_MAX
        self.saved_mass = F(0)

    def grab(self, origin, direction):
        self.index, self.distance = find_closest_vertex_to_ray(self.positions, origin, direction, self.particle_diameter)
        if self.index >= 0:
            self.grabbing = True
            self.saved_mass = F(self.inv_masses[self.index])
            self.inv_masses[self.index] = 0
        return self.index, self.distance

    def drag(self, origin, direction):
        if not self.grabbing:
            return
        o, d = np.asarray(origin, F), np.asarray(direction, F)
        mouse = o + d * self.distance
        cur = self.positions[self.index].copy()
        a = F(0.8)
        target = a * cur + (F(1) - a) * mouse
        self.positions[self.index] = target
        self.velocities[self.index] = (target - cur) / FIXED_DELTA_TIME

    def release(self):
        if self.grabbing:
            self.grabbing = False
            self.inv_masses[self.index] = self.saved_mass
