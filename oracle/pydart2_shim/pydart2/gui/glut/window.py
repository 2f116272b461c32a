"""Stub of pydart2.gui.glut.window (test infrastructure): a window that never opens."""
from pydart2.gui.opengl.scene import OpenGLScene


class GLUTWindow(object):
    def __init__(self, sim, title=None):
        self.sim = sim
        self.title = title
        self.scene = OpenGLScene()
        self.window = None
        self.window_size = (1280, 720)

    def initGL(self, w, h):
        pass

    def resizeGL(self, w, h):
        pass

    def drawGL(self):
        pass

    def mouseFunc(self, *a):
        pass

    def motionFunc(self, *a):
        pass

    def keyPressed(self, *a):
        pass

    def run(self, *a, **kw):
        pass

    def close(self):
        pass
