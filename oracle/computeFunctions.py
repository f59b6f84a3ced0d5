"""Oracle namespace with the reference's ``computeFunctions`` names (test infrastructure).

``from oracle import computeFunctions as cF`` gives every name the reference driver reaches
through ``from computeFunctions import *`` (gm:10) and ``from createPath import ...``
(gm:11), bound to the NumPy restatements.  See oracle/__init__.py for the pin status.
"""
import copy  # noqa: F401  (gm:200 uses ``copy`` re-exported from cF:1)

from .config import FDT  # noqa: F401
from .driver_ops import *  # noqa: F401,F403
from .fem import (  # noqa: F401
    bincount,
    computeConvRadBC,
    computeQuad2dFemShapeFunctions as computeQuad2dFemShapeFunctions_jax,
    computeQuad3dFemShapeFunctions as computeQuad3dFemShapeFunctions_jax,
    computeSourceFunction as computeSourceFunction_jax,
    computeSourcesL3,
    computeStateProperties,
    convert2XYZ,
    createMesh3D,
    getQuadratureCoords,
    getSampleCoords,
    solveMatrixFreeFE,
)
from .setup import (  # noqa: F401
    SetupLevels,
    SetupNonmesh,
    SetupProperties,
    calcNumNodes,
    calcStaticTmpNodesAndElements,
    calc_length_h,
    find_max_const,
    getBCindices,
    getCoarseNodesInFineRegion,
    getCoarseNodesInLargeFineRegion,
    getStaticNodesAndElements,
    getStaticSubcycle,
    getSubstrateNodes,
)
from .steppers import (  # noqa: F401
    assignBCs,
    assignBCsFine,
    computeL1Temperature,
    computeL2Temperature,
    computeSolutions,
    computeSolutions_L3,
    jit_constrain_v,
    levelMaxMin,
    melting_temp,
    moveEverything,
    move_fine_mesh,
    stepGOMELT,
    stepGOMELTDwellTime,
    subcycleGOMELT,
    substitute_Tbar,
    update_overlap_nodes_coords,
    update_overlap_nodes_coords_L1L2,
    update_overlap_nodes_coords_L2,
    updateStateProperties,
)
from .toolpath import count_lines, format_fixed, parsingGcode  # noqa: F401
from .transfer import (  # noqa: F401
    compute3DN,
    computeCoarseFineShapeFunctions,
    computeCoarseTprimeMassTerm as computeCoarseTprimeMassTerm_jax,
    computeCoarseTprimeTerm as computeCoarseTprimeTerm_jax,
    computeL1TprimeTerms_Part1,
    computeL1TprimeTerms_Part2,
    computeL2TprimeTerms_Part1,
    computeL2TprimeTerms_Part2,
    computeLevelSource,
    computeSources,
    getBothNewTprimes,
    getNewTprime,
    getOverlapRegion,
    interpolate_w_matrix,
    interpolatePoints,
    interpolatePointsMatrix,
)
