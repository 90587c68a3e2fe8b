# Stand-in package so `from isaacgym.torch_utils import *` resolves when the UNMODIFIED reference
# is imported in the build container (Isaac Gym itself is not installable here).  Test
# infrastructure only: nothing under pulse_b200/ imports this.
