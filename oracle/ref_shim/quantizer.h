/* Shim used ONLY while compiling the reference's dlquant/quantizer.c (from where it lies under
 * /root/reference) into oracle/_ref/.  It pulls in the reference header unmodified, then restores the
 * MSVC LLP64 type widths the DLL was built for: `long` is 32-bit under MSVC x64 but 64-bit under gcc/Linux
 * (quantizer.h:12-14 define slong/ulong as `signed long`/`unsigned long`). */
#include TM_REF_QUANTIZER_H
#undef slong
#undef ulong
#define slong signed int
#define ulong unsigned int
