"""Stub of pydart2.gui.trackball (test infrastructure)."""


class Trackball(object):
    def __init__(self, theta=0.0, phi=0.0, zoom=1.0, **kw):
        self.trans = [0.0, 0.0, 0.0]

    def _set_theta(self, theta):
        self.theta = theta

    def _set_phi(self, phi):
        self.phi = phi
