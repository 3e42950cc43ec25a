-- ARAP image-deformation energy, in the Opt problem-specification language.
--
-- libarapb200 does not interpret this file: Opt_ProblemDefine only checks that it is present and
-- declares the ARAP problem (include/Opt.h), because the derivatives of exactly this energy are
-- hand-derived in arap_flow_b200/csrc/grid_math.cuh.  It is kept so that the reference's drivers,
-- which require $ARAP_PLAN to exist (ARAP/deformation/src/main.cpp:206-213), run unchanged, and it
-- states the problem in a form the original Opt would also accept.
--
--   unknowns   Offset(W,H) : float2   absolute deformed position of every pixel   (parameter 0)
--              Angle(W,H)  : float    per-pixel rotation                          (parameter 1)
--   inputs     UrShape (2), Constraints (3; (-1,-1) = unconstrained), Mask (4; 0 = deformable)
--   weights    w_fitSqrt (5), w_regSqrt (6)
local W, H        = Dim("W", 0), Dim("H", 1)
local Offset      = Unknown("Offset", opt_float2, { W, H }, 0)
local Angle       = Unknown("Angle", opt_float, { W, H }, 1)
local UrShape     = Array("UrShape", opt_float2, { W, H }, 2)
local Constraints = Array("Constraints", opt_float2, { W, H }, 3)
local Mask        = Array("Mask", opt_float, { W, H }, 4)
local w_fitSqrt   = Param("w_fitSqrt", float, 5)
local w_regSqrt   = Param("w_regSqrt", float, 6)

UsePreconditioner(true)
-- pixels off the object are not unknowns at all
Exclude(Not(eq(Mask(0, 0), 0)))

-- rigidity: each of the four edges of a pixel should be the rest-pose edge rotated by the pixel's angle
local neighbours = { { 1, 0 }, { -1, 0 }, { 0, 1 }, { 0, -1 } }
for dx, dy in Stencil(neighbours) do
    local edge_now  = Offset(0, 0) - Offset(dx, dy)
    local edge_rest = UrShape(0, 0) - UrShape(dx, dy)
    local both_on_object = InBounds(dx, dy) * eq(Mask(dx, dy), 0) * eq(Mask(0, 0), 0)
    Energy(Select(both_on_object, w_regSqrt * (edge_now - Rotate2D(Angle(0, 0), edge_rest)), 0))
end

-- fitting: constrained pixels should land on their (continuation-interpolated) match target
local has_target = All(greatereq(Constraints(0, 0), 0))
Energy(w_fitSqrt * Select(has_target, Offset(0, 0) - Constraints(0, 0), 0.0))
