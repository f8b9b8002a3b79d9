// Placeholder: until the hand-derived stable-Neo-Hookean kernel lands the tet potentials use the AD evaluation.
#pragma once
namespace sb {
template<bool COMPLETE> static void launch_tet_analytic_pgh(const EvalArgs& a, cudaStream_t s)
{
    launch_pgh<sbpot::EnergyTetStrainT<COMPLETE>>(a, s);
}
}  // namespace sb
