"""fsilbm3d_b200 -- host-side mirror of the FSILBM3D hot-path interface over libfsilbm_b200.so.

The product is the C-ABI shared library (include/fsilbm.h, csrc/*.cu, sm_100a).  These Python
classes carry the reference's names (LBMBlock, VirtualBody, tree_collision_streaming_IBM_FEM) so a
user of the reference finds the same call points; they only marshal arguments into the C ABI.
There is no CPU path: importing works anywhere, but every compute call raises FsilbmError when
the CUDA library or a B200 is missing.
"""
from .flow_condition import FlowCondType
from ._lib import FsilbmError, lib, library_path, exported_symbols, declared_symbols
from .fluid_domain import LBMBlock
from .solid_body import VirtualBody, RigidPlate
from . import flow_io
from .block_comm import (tree_collision_streaming_IBM_FEM, slab_range, init_process_group, halo_plan, CommPair, RemoteSon, blockTreeNode,
                         build_block_tree, CompareBlocks, find_carrier_fluidblock, set_blktime_all, ibm_box_participants, son_slab_plan)

__all__ = [
    "FlowCondType", "FsilbmError", "lib", "library_path", "exported_symbols", "declared_symbols",
    "LBMBlock", "VirtualBody", "RigidPlate", "flow_io", "tree_collision_streaming_IBM_FEM", "slab_range",
    "init_process_group", "halo_plan", "CommPair", "RemoteSon", "blockTreeNode", "build_block_tree", "CompareBlocks",
    "find_carrier_fluidblock", "set_blktime_all", "ibm_box_participants", "son_slab_plan",
]
