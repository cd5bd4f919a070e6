"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  Nothing under astro-sph-tools_b200/ imports it: the product path fails loudly when the CUDA
library is missing, it never routes through here.
"""
from .oracle import *  # noqa: F401,F403
