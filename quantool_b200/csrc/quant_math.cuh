// compressed-tensors quantization primitives as device functions:
//   calculate_qparams  CT/quantization/utils/helpers.py:50-137
//   _quantize / _dequantize / _quantize_dequantize  CT/quantization/lifecycle/forward_helpers.py:176-268
//   round_to_quantized_type_args (clamp, then torch.round = half-to-even)  CT/quantization/quant_args.py:439-475
#pragma once
#include "common.cuh"

namespace qt {

struct QRange { float qmin, qmax; };
QT_HD QRange int_range(int num_bits) {
    const float half = (float)(1 << (num_bits - 1));
    return QRange{-half, half - 1.f};
}

// min/max -> (scale, zero_point).  zero_point is returned as a float holding an integer value.
QT_D void calc_qparams(float mn, float mx, int num_bits, bool symmetric, float& scale, float& zp) {
    mn = fminf(mn, 0.f);
    mx = fmaxf(mx, 0.f);
    const QRange r = int_range(num_bits);
    const float bit_range = r.qmax - r.qmin;
    if (symmetric) {
        const float max_val_pos = fmaxf(fabsf(mn), fabsf(mx));
        scale = max_val_pos / (bit_range / 2.f);
        zp = 0.f;
    } else {
        scale = (mx - mn) / bit_range;
        float z = r.qmin - (mn / scale);            // scale == 0 -> NaN
        z = fminf(fmaxf(z, r.qmin), r.qmax);
        z = fminf(fmaxf(z, -128.f), 127.f);         // zp_dtype = int8
        z = rintf(z);
        zp = (z == z) ? z : 0.f;                     // NaN -> int8 cast gives 0
    }
    if (scale == 0.f) scale = 1.1920928955078125e-07f;  // torch.finfo(float32).eps
}

// fp32 fake-quantize: clamp(round(x/scale + zp)) then (q - zp) * scale
QT_D float fake_quant(float x, float scale, float zp, QRange r, float& q_out) {
    float v = x / scale;
    v = v + zp;
    v = fminf(fmaxf(v, r.qmin), r.qmax);
    v = rintf(v);
    q_out = v;
    return (v - zp) * scale;
}

}  // namespace qt
