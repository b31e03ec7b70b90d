// Package baseline times the REAL go-muse (github.com/aouyang1/go-muse, the unmodified reference) on the
// shapes bench.py uses, so that anyone with a Go toolchain can put the true host-CPU number beside the B200
// figures.  It is shipped unexecuted: this repository's image has no Go (BASELINE.md section 2); the CPU arm
// bench.py reports is oracle/muse_oracle.c, a C restatement of the same path.
//
//	cd baseline/go && go mod init musebaseline && go get github.com/aouyang1/go-muse@master && \
//	  GOMAXPROCS=$(nproc) go test -vet=off -run xxx -bench . -benchtime 3x
//
// series-samples/s = S*N / (ns/op * 1e-9); S and N are printed by each benchmark.
// MUSE_BASELINE_SERIES overrides the C3 slice size (default 100000 of the 1,000,000 series).
package baseline

import (
	"math/rand"
	"os"
	"runtime"
	"strconv"
	"testing"

	muse "github.com/aouyang1/go-muse"
)

// siggen-style rows of SURVEY.md section 8d: kind = i mod 3 -> rect+noise, line+noise, noise.
func row(rng *rand.Rand, i, n int) []float64 {
	y := make([]float64, n)
	for t := range y {
		y[t] = 0.1 * (rng.Float64() - 0.5)
	}
	switch i % 3 {
	case 0:
		amp := 0.5 + 39.5*rng.Float64()
		mid := n/2 - n/8 + rng.Intn(n/4)
		w := 3 + rng.Intn(18)
		for t := mid - w/2; t < mid-w/2+w && t < n; t++ {
			if t >= 0 {
				y[t] += amp
			}
		}
	case 1:
		slope := 0.02 * (rng.Float64() - 0.5)
		for t := range y {
			y[t] += slope * float64(t)
		}
	}
	return y
}

func reference(rng *rand.Rand, n int) []float64 {
	y := make([]float64, n)
	for t := range y {
		y[t] = 0.1 * (rng.Float64() - 0.5)
	}
	for t := n/2 - 5; t < n/2+5; t++ {
		y[t] += 1.5
	}
	return y
}

func run(b *testing.B, series, n, maxLag, topN int, threshold float64, groupBy []string) {
	rng := rand.New(rand.NewSource(20261018))
	comp := muse.NewGroup("comparison")
	for i := 0; i < series; i++ {
		lm := muse.LabelMap{"graph": "g" + strconv.Itoa(i/1000), "host": "h" + strconv.Itoa(i%1000)}
		if err := comp.Add(muse.NewSeries(row(rng, i, n), muse.NewLabels(lm))); err != nil {
			b.Fatal(err)
		}
	}
	ref := muse.NewSeries(reference(rng, n), muse.NewLabels(muse.LabelMap{"graph": "ref"}))
	cc := runtime.GOMAXPROCS(0)
	b.ResetTimer()
	for it := 0; it < b.N; it++ {
		batch, err := muse.NewBatch(ref, comp, muse.NewResults(maxLag, topN, threshold, muse.SignFilter_ANY), cc)
		if err != nil {
			b.Fatal(err)
		}
		if err := batch.Run(groupBy); err != nil {
			b.Fatal(err)
		}
		batch.Results.Fetch()
	}
	b.StopTimer()
	sec := b.Elapsed().Seconds() / float64(b.N)
	b.ReportMetric(float64(series)*float64(n)/sec, "series-samples/s")
	b.Logf("S=%d N=%d maxLag=%d topN=%d threshold=%g groupBy=%v GOMAXPROCS=%d", series, n, maxLag, topN, threshold, groupBy, cc)
}

// BASELINE.json configs[1]: BenchmarkMuseBatchRunLarge's shape (muse_batch_test.go:137-164).
func BenchmarkC2(b *testing.B) { run(b, 5000, 480, 10, 20, 0, []string{"graph"}) }

// BASELINE.json configs[2]: a slice of the 1 M x 1440 store, ungrouped, maxLag 60, topN 100, threshold 0.5.
func BenchmarkC3Slice(b *testing.B) {
	s := 100000
	if v, err := strconv.Atoi(os.Getenv("MUSE_BASELINE_SERIES")); err == nil && v > 0 {
		s = v
	}
	run(b, s, 1440, 60, 100, 0.5, nil)
}
