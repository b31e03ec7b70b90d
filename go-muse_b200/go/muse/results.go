package muse

import (
	"container/heap"
	"math"
	"sync"
)

// Results tracks the top scores given a maximum lag, top N and score threshold (go-muse results.go:11-87).
type Results struct {
	sync.Mutex
	MaxLag     int
	TopN       int
	Threshold  float64
	SignFilter SignFilter
	scores     Scores
}

type SignFilter int

const (
	SignFilter_POS = 1
	SignFilter_NEG = -1
	SignFilter_ANY = 0
)

// NewResults creates a new instance of results to track the top similar graphs.
func NewResults(maxLag int, topN int, threshold float64, signFilter SignFilter) *Results {
	scores := make(Scores, 0, topN)
	heap.Init(&scores)
	return &Results{MaxLag: maxLag, TopN: topN, Threshold: threshold, SignFilter: signFilter, scores: scores}
}

func (r *Results) passed(s Score) bool {
	return math.Abs(float64(s.Lag)) <= float64(r.MaxLag) &&
		math.Abs(s.PercentScore) >= r.Threshold &&
		(r.SignFilter == SignFilter_ANY ||
			(s.PercentScore > 0 && r.SignFilter == SignFilter_POS) ||
			(s.PercentScore < 0 && r.SignFilter == SignFilter_NEG))
}

// Update records the input score.
func (r *Results) Update(s Score) {
	if s.Labels == nil {
		return
	}
	r.Lock()
	if r.passed(s) {
		if r.scores.Len() == r.TopN {
			if r.TopN > 0 && math.Abs(s.PercentScore) > math.Abs(r.scores[0].PercentScore) {
				heap.Pop(&r.scores)
				heap.Push(&r.scores, s)
			}
		} else {
			heap.Push(&r.scores, s)
		}
	}
	r.Unlock()
}

// Fetch drains the heap and returns the scores in descending |score| order with their mean.
func (r *Results) Fetch() (Scores, float64) {
	s := make(Scores, len(r.scores))
	var sum float64
	n := len(r.scores)
	for i := n - 1; i >= 0; i-- {
		sc := heap.Pop(&r.scores).(Score)
		sum += math.Abs(sc.PercentScore)
		s[i] = sc
	}
	return s, sum / float64(n)
}
