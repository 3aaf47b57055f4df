/* ORACLE — TEST INFRASTRUCTURE ONLY. Op codes of the type-erased oracle C API (oracle_capi.cc).
 * The numbering is private to the oracle; the product's C ABI (include/gadcu.h) has named entry
 * points instead. */
#pragma once

enum oracle_op_code {
  OOP_VARIABLE = 0,
  OOP_CONSTANT = 1,
  OOP_ADD = 2,
  OOP_ADD_ALL = 3,
  OOP_SUB = 4,
  OOP_MUL = 5,
  OOP_ZEROS = 6,
  OOP_ONES = 7,
  OOP_NEG = 8,
  OOP_SETC = 9,   /* sp[0]=c, sp[1]=1 if the constant is an integer (i16) */
  OOP_ADDC = 10,
  OOP_MULC = 11,
  OOP_POWC = 12,
  OOP_EXP = 13,
  OOP_LOG = 14,
  OOP_LOG1P = 15,
  OOP_SIN = 16,
  OOP_COS = 17,
  OOP_TANH = 18,
  OOP_SIGMOID = 19,
  OOP_RECIPROCAL = 20,
  OOP_SQRT = 21,
  OOP_DIV = 22,
  OOP_POW = 23,
  OOP_SELECT_ARGMAX = 24, /* ins: v0, v1, r0|NULL, r1|NULL (n_in = 4) */
  OOP_MIN = 25,
  OOP_MAX = 26,
  OOP_ABS = 27,
  OOP_SIGN = 28,
  OOP_RELU = 29,
  OOP_FLAT = 30,
  OOP_MODDIMS = 31,  /* dp[0..4) */
  OOP_TILE_AS = 32,  /* dp[0..4) */
  OOP_SUM_AS = 33,   /* dp[0..4) */
  OOP_CONSTANT_AS = 34, /* ins: scalar; dp[0..4) */
  OOP_AS_SCALAR = 35,
  OOP_SCALE = 36,    /* ins: scalar, array */
  OOP_DOT = 37,
  OOP_NORM2 = 38,
  OOP_MATMUL = 39,   /* sp[0..4) = transposed1, conjugated1, transposed2, conjugated2 */
  OOP_TRANSPOSE = 40, /* sp[0] = conjugate */
  OOP_MATMUL_NN = 41,
  OOP_MAX_AS = 42,   /* dp[0..4) */
  OOP_ARGMAX_AS = 43,
  OOP_SOFTMAX_AS = 44,
  OOP__COUNT = 45
};

enum oracle_alg_kind { OALG_CHECK = 0, OALG_EVAL = 1, OALG_GRAPH1 = 2, OALG_GRAPHN = 3 };
