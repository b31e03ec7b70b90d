package muse

import "math"

// Scores is a slice of individual Score and implements heap.Interface on |PercentScore|.
type Scores []Score

// Score keeps track of the cross correlation score and the related series (go-muse scores.go:11-15).
type Score struct {
	Labels       *Labels `json:"labels"`
	Lag          int     `json:"lag"`
	PercentScore float64 `json:"percentScore"`
}

func (s Scores) Len() int            { return len(s) }
func (s Scores) Swap(i, j int)       { s[i], s[j] = s[j], s[i] }
func (s Scores) Less(i, j int) bool  { return math.Abs(s[i].PercentScore) < math.Abs(s[j].PercentScore) }
func (s *Scores) Push(x interface{}) { *s = append(*s, x.(Score)) }
func (s *Scores) Pop() interface{} {
	x := (*s)[len(*s)-1]
	*s = (*s)[:len(*s)-1]
	return x
}
