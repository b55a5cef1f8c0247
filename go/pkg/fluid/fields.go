package fluid

import "fmt"

// ScalarField keeps the reference's type (scalar_field.go:7-22).
type ScalarField struct {
	NumX, NumY         int
	values             []float32
	MinValue, MaxValue float32
}

func (s ScalarField) Value(i, j int) (float32, error) {
	if i < 0 || i >= s.NumX {
		return 0.0, fmt.Errorf("x index (%d) out of range, must be between 0 and %d", i, s.NumX-1)
	}
	if j < 0 || j >= s.NumY {
		return 0.0, fmt.Errorf("y index (%d) out of range, must be between 0 and %d", j, s.NumY-1)
	}
	return s.values[i*s.NumY+j], nil
}

// VectorField keeps the reference's type (vector_field.go:5-19).
type VectorField struct {
	NumX, NumY       int
	valuesU, valuesV []float32
}

func (v VectorField) Value(i, j int) (float32, float32, error) {
	if i < 0 || i >= v.NumX {
		return 0.0, 0.0, fmt.Errorf("x index out of range, must be between 0 and %d", v.NumX-1)
	}
	if j < 0 || j >= v.NumY {
		return 0.0, 0.0, fmt.Errorf("y index out of range, must be between 0 and %d", v.NumY-1)
	}
	return v.valuesU[i*v.NumY+j], v.valuesV[i*v.NumY+j], nil
}
