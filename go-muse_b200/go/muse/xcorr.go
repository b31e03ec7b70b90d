package muse

/*
#include "muse_b200.h"
*/
import "C"

import "unsafe"

// xCorr computes the cross correlation slice between x and y, the lag of the maximum absolute value
// and that value (go-muse xcorr.go:102-153), for any n -- on the device (muse_xcorr: the circular
// correlation evaluated directly up to 4096 lags, by FFT passes above).  n is raised to max(n, len(x), len(y)); with normalize both
// inputs are z-normalized first and a constant input yields (nil, 0, 0) exactly as the reference does.
func xCorr(x []float64, y []float64, n int, normalize bool) ([]float64, int, float64) {
	if len(x) == 0 || len(y) == 0 {
		return nil, 0, 0
	}
	c, err := deviceContext()
	if err != nil {
		return nil, 0, 0
	}
	nn := n
	if len(x) > nn {
		nn = len(x)
	}
	if len(y) > nn {
		nn = len(y)
	}
	cc := make([]float64, nn)
	var nOut, lag C.int64_t
	var val C.double
	var stdZero C.int32_t
	norm := C.int32_t(0)
	if normalize {
		norm = 1
	}
	rc := C.muse_xcorr(c, (*C.double)(unsafe.Pointer(&x[0])), C.int64_t(len(x)),
		(*C.double)(unsafe.Pointer(&y[0])), C.int64_t(len(y)), C.int64_t(n), norm,
		(*C.double)(unsafe.Pointer(&cc[0])), C.int64_t(len(cc)), &nOut, &lag, &val, &stdZero)
	if rc != C.MUSE_OK || stdZero != 0 {
		return nil, 0, 0
	}
	return cc[:int(nOut)], int(lag), float64(val)
}
