/* ref_root_shim.c — C-callable entry onto the reference's root adm.c (+ root gaussj.c, nrutil.c),
 * compiled in place from /root/reference into oracle/_ref/.  TEST INFRASTRUCTURE ONLY. */
#include <stdlib.h>
double *adm(double *x, int n, int *check, void (*funcvmix)(int, double *, double *), int flag); /* adm.c:24-27 */

int ref_adm(void (*f)(int, double *, double *), double *x, int n) {
  int check = 1;
  double *xnew = adm(x, n, &check, f, 0); /* flag 0: 0-based arrays (adm.c:17) */
  (void)xnew; /* adm.c:8-10 frees xnew's siblings and returns a dangling-free pointer we do not own */
  return check;
}
