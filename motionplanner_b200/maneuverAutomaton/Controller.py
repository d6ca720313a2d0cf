"""Controller object returned next to the planned trajectory.

Same attributes and call as reference maneuverAutomaton/Controller.py:90-120 (``FeedbackController``): attribute names
``x_ref, u_ref, time, K`` and ``get_control_input(time, x) -> u``.  Only the controller of the time-point automaton
is carried; control laws are outside the collision-check path."""
import numpy as np


class FeedbackController:
    """tracks a reference trajectory: u = u_ref + K (x - x_ref(t)), x_ref linearly interpolated between time points"""

    def __init__(self, x_ref, u_ref, time, K):
        self.x_ref = x_ref      # reference trajectory
        self.u_ref = u_ref      # inputs of the reference trajectory
        self.time = time        # time points of the reference trajectory
        self.K = K              # one feedback matrix per segment

    def get_control_input(self, time, x):
        # the orientation error is taken on the circle
        err = x[3, 0] - self.x_ref[3, 0]
        x[3, 0] = self.x_ref[3, 0] + np.mod(err + np.pi, 2 * np.pi) - np.pi
        ind = int(np.nonzero(np.asarray(self.time) <= time)[0][-1])
        w = (time - self.time[ind]) / (self.time[ind + 1] - self.time[ind])
        x_ref = self.x_ref[:, [ind]] + w * (self.x_ref[:, [ind + 1]] - self.x_ref[:, [ind]])
        u = self.u_ref[:, [ind]] + self.K[ind] @ (x - x_ref)
        return u[:, 0]
