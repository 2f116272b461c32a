"""Stub of pydart2.gui.opengl.scene (test infrastructure)."""
from pydart2.gui.trackball import Trackball


class OpenGLScene(object):
    def __init__(self, *a, **kw):
        self.cameras = [Trackball()]
        self.tb = self.cameras[0]

    def add_camera(self, cam, name=None):
        self.cameras.append(cam)

    def num_cameras(self):
        return len(self.cameras)

    def set_camera(self, i):
        self.tb = self.cameras[i]

    def render(self, sim):
        pass
