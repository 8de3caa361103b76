// Element-wise pieces shared by the multi-kernel finish (finish.cu) and the single-CTA finish of short signals
// (critic_stats.cu): one definition, so both paths round the same way.
#pragma once
#include <stdint.h>

namespace hypad {

// combine_scores (utils/anomaly_detection_utils.py:336-362) and score_anomalies' "sum" (:554-570) for one position
template <typename TR>
__device__ __forceinline__ double combine_value(int mode, double cv, TR r, double uv, double lambda_rec) {
    const double rv = (double)r;
    switch (mode) {
        case 0: return cv * rv;                                    // mult
        case 1: return (cv * rv) * uv;                             // uncertainty
        case 2: return 0.2 * cv + 0.8 * rv;                        // sum
        case 3: return cv;                                         // critic
        case 4: return cv * uv;                                    // critic_uncertainty
        case 5: return (0.5 * cv) * uv + (0.5 * rv) * uv;          // sum_uncertainty
        case 6: return rv;                                         // rec
        case 7:                                                    // rec_uncertainty: an fp32 tensor times the fp32 norms stays fp32
            if (sizeof(TR) == 4) return (double)__fmul_rn((float)rv, (float)uv);
            return rv * uv;
        default: return (1.0 - lambda_rec) * (cv - 1.0) + lambda_rec * (rv - 1.0);  // score_anomalies "sum"
    }
}

// the critic z-score |x - mean_band| / std + 1 (:322-325)
__device__ __forceinline__ double critic_z(double x, double mu, double sd) { return fabs((x - mu) / sd) + 1.0; }

}  // namespace hypad
