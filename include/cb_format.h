/* cb_format.h -- the fixed-precision midpoint-radius ball format used in device memory.
 *
 * This replaces, on the device, Arb's heap-allocated acb_struct (arf_struct midpoint +
 * mag_struct radius per real part) that the reference stores everywhere:
 *   reference src/acb_ode.h:9-15 (acb_ptr polys), src/acb_ode.h:45-50 (acb_poly_struct *gens).
 *
 * One REAL ball occupies one "slot".  A complex ball is two adjacent slots (re, im).
 * A batch of S slots at precision prec = 32*L bits is stored struct-of-arrays:
 *
 *   mant : uint32_t[L][S]   limb-major ("limb k of slot s" at mant[k*S + s]); little-endian
 *                           limbs; normalised (bit 31 of limb L-1 set) unless the midpoint is 0.
 *                           A warp reading limb k of 32 consecutive slots touches one 128-byte line.
 *   meta : cb_meta[S]       16 bytes per slot, read as one uint4.
 *
 * Midpoint value  = (-1)^sign * (M / 2^(32 L)) * 2^exp,   M = sum mant[k] 2^(32 k)  in [2^(32L-1), 2^(32L)).
 * Radius value    = rman * 2^(rexp - 30),  rman in [2^29, 2^30), or rman == 0 (exact), or
 *                   rman == CB_MAG_INF_MAN (infinite radius).
 *
 * Bytes per real ball: 4 L + 16  (144 / 272 / 528 at 1024 / 2048 / 4096 bits).
 *
 * A series batch of N coefficients for B complex instances is [N][L][2B] mantissa words and
 * [N][2B] meta words; slot index inside one coefficient = 2*instance + part (0 = re, 1 = im).
 */
#ifndef CB_FORMAT_H
#define CB_FORMAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_FLAG_SIGN 1u /* midpoint negative */
#define CB_FLAG_ZERO 2u /* midpoint exactly zero (mant all zero, exp 0) */
#define CB_FLAG_NAN  4u /* indeterminate: midpoint NaN (Arb: acb_indeterminate) */

#define CB_MAG_BITS 30
#define CB_MAG_INF_MAN 0xFFFFFFFFu

/* Midpoint exponents outside +-CB_EXP_LIMIT turn the ball indeterminate (Arb has unbounded
 * fmpz exponents; the device keeps 32-bit ones). Radius exponents saturate, see cb_mag.h. */
#define CB_EXP_LIMIT (1 << 28)

typedef struct {
    int32_t exp;    /* midpoint exponent (0 if zero) */
    uint32_t flags; /* CB_FLAG_* */
    uint32_t rman;  /* radius mantissa (30 bits), 0, or CB_MAG_INF_MAN */
    int32_t rexp;   /* radius exponent */
} cb_meta;

#define CB_MAX_LIMBS 128 /* 4096 bits */

#ifdef __cplusplus
}
#endif
#endif
