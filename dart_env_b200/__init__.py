"""B200-native batched DART stepper for DartEnv (drop-in for the pydart2 World.step() path)."""
__version__ = "0.1.0"
