// CPU-side check of the folding arithmetic in mahout_b200/csrc/cm_hash.cuh (the header compiles for the
// host for exactly this purpose): reads "a b key w" lines on stdin, prints the column of
// ((a*key + b) mod (2^63 - 25)) mod w through cmh_residue / cmh_column.  Arithmetic only -- the product
// path runs on the GPU.
#include <stdio.h>

#include "../../mahout_b200/csrc/cm_hash.cuh"

int main() {
  long long a, b, key;
  unsigned w;
  while (scanf("%lld %lld %lld %u", &a, &b, &key, &w) == 4) {
    const unsigned wmask = (w > 1 && (w & (w - 1)) == 0) ? w - 1 : 0u;
    printf("%u\n", cmh_column(cmh_residue(a), cmh_residue(b), cmh_residue(key), w, wmask));
  }
  return 0;
}
