// weno.cuh -- fifth-order WENO weights and interpolation (device functions).
// Reference: include/interpolation.h:265-500 (_WENOWeights_v_JS_/_M_/_Z_/_YC_),
// src/InterpolationFunctions/Interp1PrimFifthOrderWENO.c:148-158.
#pragma once
#include "hpb_internal.h"

#ifndef HPB_DEV
#define HPB_DEV __device__ __forceinline__
#endif

// smoothness indicators of the stencil (m3,m2,m1,p1,p2)
HPB_DEV void weno_betas(double m3, double m2, double m1, double p1, double p2, double& b1, double& b2, double& b3)
{
  const double thirteen_by_twelve = 13.0 / 12.0, one_fourth = 0.25;
  const double d1 = m3 - 2 * m2 + m1, e1 = m3 - 4 * m2 + 3 * m1;
  const double d2 = m2 - 2 * m1 + p1, e2 = m2 - p1;
  const double d3 = m1 - 2 * p1 + p2, e3 = 3 * m1 - 4 * p1 + p2;
  b1 = thirteen_by_twelve * d1 * d1 + one_fourth * e1 * e1;
  b2 = thirteen_by_twelve * d2 * d2 + one_fourth * e2 * e2;
  b3 = thirteen_by_twelve * d3 * d3 + one_fourth * e3 * e3;
}

// One scalar weight set in the reference's own form (used by the generic path and the
// fine-grained API) for the optimal weights (c1, c2, c3), p = 2.
HPB_DEV void weno_weights_ref_c(int type, double eps, double c1, double c2, double c3,
                                double m3, double m2, double m1, double p1, double p2,
                                double& w1, double& w2, double& w3)
{
  double b1, b2, b3, a1, a2, a3;
  weno_betas(m3, m2, m1, p1, p2, b1, b2, b3);
  if (type == HPB_WENO_JS || type == HPB_WENO_M) {
    a1 = c1 / ((b1 + eps) * (b1 + eps));
    a2 = c2 / ((b2 + eps) * (b2 + eps));
    a3 = c3 / ((b3 + eps) * (b3 + eps));
  } else {
    double tau;
    if (type == HPB_WENO_Z) tau = fabs(b3 - b1);
    else { const double t = m3 - 4 * m2 + 6 * m1 - 4 * p1 + p2; tau = t * t; }
    const double r1 = tau / (b1 + eps), r2 = tau / (b2 + eps), r3 = tau / (b3 + eps);
    a1 = c1 * (1.0 + r1 * r1);
    a2 = c2 * (1.0 + r2 * r2);
    a3 = c3 * (1.0 + r3 * r3);
  }
  double a_sum_inv = 1.0 / (a1 + a2 + a3);
  w1 = a1 * a_sum_inv; w2 = a2 * a_sum_inv; w3 = a3 * a_sum_inv;
  if (type == HPB_WENO_M) {
    a1 = w1 * (c1 + c1 * c1 - 3 * c1 * w1 + w1 * w1) / (c1 * c1 + w1 * (1.0 - 2.0 * c1));
    a2 = w2 * (c2 + c2 * c2 - 3 * c2 * w2 + w2 * w2) / (c2 * c2 + w2 * (1.0 - 2.0 * c2));
    a3 = w3 * (c3 + c3 * c3 - 3 * c3 * w3 + w3 * w3) / (c3 * c3 + w3 * (1.0 - 2.0 * c3));
    a_sum_inv = 1.0 / (a1 + a2 + a3);
    w1 = a1 * a_sum_inv; w2 = a2 * a_sum_inv; w3 = a3 * a_sum_inv;
  }
}

// WENO5: optimal weights (0.1, 0.6, 0.3) (interpolation.h:228-232)
HPB_DEV void weno_weights_ref(int type, double eps, double m3, double m2, double m1, double p1, double p2,
                              double& w1, double& w2, double& w3)
{
  weno_weights_ref_c(type, eps, 0.1, 0.6, 0.3, m3, m2, m1, p1, p2, w1, w2, w3);
}

// fifth-order interpolant from the three candidate stencils
HPB_DEV double weno_combine(double w1, double w2, double w3, double m3, double m2, double m1, double p1, double p2)
{
  const double one_sixth = 1.0 / 6.0;
  const double f1 = (2 * one_sixth) * m3 + (-7 * one_sixth) * m2 + (11 * one_sixth) * m1;
  const double f2 = (-one_sixth) * m2 + (5 * one_sixth) * m1 + (2 * one_sixth) * p1;
  const double f3 = (2 * one_sixth) * m1 + (5 * one_sixth) * p1 + (-one_sixth) * p2;
  return w1 * f1 + w2 * f2 + w3 * f3;
}
