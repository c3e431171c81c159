// XnaStub.cs -- the three Microsoft.Xna.Framework types the reference's physics sources use (Vector2, Matrix, Color), with the
// arithmetic of MonoGame.Framework 3.8 written out operation for operation (SURVEY.md Appendix C lists the call sites):
// reciprocal-multiply division, Normalize through 1f / sqrt, CreateRotationZ through double Math.Cos / Math.Sin, no FMA.
// If the real MonoGame.Framework assembly is available, delete this file and reference the package instead: the output of
// the harness must not change.
using System;

namespace Microsoft.Xna.Framework;

public struct Vector2 : IEquatable<Vector2>
{
    public float X;
    public float Y;

    public Vector2(float x, float y)
    {
        X = x;
        Y = y;
    }

    public Vector2(float value)
    {
        X = value;
        Y = value;
    }

    public static Vector2 Zero => new Vector2(0f, 0f);
    public static Vector2 One => new Vector2(1f, 1f);

    public static Vector2 operator -(Vector2 value)
    {
        value.X = -value.X;
        value.Y = -value.Y;
        return value;
    }

    public static Vector2 operator +(Vector2 value1, Vector2 value2)
    {
        value1.X += value2.X;
        value1.Y += value2.Y;
        return value1;
    }

    public static Vector2 operator -(Vector2 value1, Vector2 value2)
    {
        value1.X -= value2.X;
        value1.Y -= value2.Y;
        return value1;
    }

    public static Vector2 operator *(Vector2 value1, Vector2 value2)
    {
        value1.X *= value2.X;
        value1.Y *= value2.Y;
        return value1;
    }

    public static Vector2 operator *(Vector2 value, float scaleFactor)
    {
        value.X *= scaleFactor;
        value.Y *= scaleFactor;
        return value;
    }

    public static Vector2 operator *(float scaleFactor, Vector2 value)
    {
        value.X *= scaleFactor;
        value.Y *= scaleFactor;
        return value;
    }

    public static Vector2 operator /(Vector2 value1, float divider)
    {
        float factor = 1 / divider;
        value1.X *= factor;
        value1.Y *= factor;
        return value1;
    }

    public static Vector2 Divide(Vector2 value1, float divider)
    {
        float factor = 1 / divider;
        value1.X *= factor;
        value1.Y *= factor;
        return value1;
    }

    public static bool operator ==(Vector2 value1, Vector2 value2) => value1.X == value2.X && value1.Y == value2.Y;

    public static bool operator !=(Vector2 value1, Vector2 value2) => value1.X != value2.X || value1.Y != value2.Y;

    public static float Dot(Vector2 value1, Vector2 value2) => (value1.X * value2.X) + (value1.Y * value2.Y);

    public float Length() => MathF.Sqrt((X * X) + (Y * Y));

    public float LengthSquared() => (X * X) + (Y * Y);

    public void Normalize()
    {
        float val = 1.0f / MathF.Sqrt((X * X) + (Y * Y));
        X *= val;
        Y *= val;
    }

    public static Vector2 Transform(Vector2 position, Matrix matrix) =>
        new Vector2((position.X * matrix.M11) + (position.Y * matrix.M21) + matrix.M41,
                    (position.X * matrix.M12) + (position.Y * matrix.M22) + matrix.M42);

    public bool Equals(Vector2 other) => X == other.X && Y == other.Y;   // List<Vector2>.Remove / Contains use this (ContactPoints.cs:47-50)

    public override bool Equals(object obj) => obj is Vector2 v && Equals(v);

    public override int GetHashCode() => HashCode.Combine(X, Y);

    public override string ToString() => "{X:" + X + " Y:" + Y + "}";
}

public struct Matrix
{
    public float M11, M12, M13, M14;
    public float M21, M22, M23, M24;
    public float M31, M32, M33, M34;
    public float M41, M42, M43, M44;

    public static Matrix Identity
    {
        get
        {
            Matrix m = default;
            m.M11 = 1f;
            m.M22 = 1f;
            m.M33 = 1f;
            m.M44 = 1f;
            return m;
        }
    }

    public static Matrix CreateRotationZ(float radians)
    {
        Matrix result = Identity;
        var val1 = (float)Math.Cos(radians);
        var val2 = (float)Math.Sin(radians);
        result.M11 = val1;
        result.M12 = val2;
        result.M21 = -val2;
        result.M22 = val1;
        return result;
    }
}

public struct Color
{
    public byte R, G, B, A;

    public Color(int r, int g, int b, int a = 255)
    {
        R = (byte)r;
        G = (byte)g;
        B = (byte)b;
        A = (byte)a;
    }

    // the named colours the Materials use (Materials/*.cs:10); the physics never reads them
    public static Color White => new Color(255, 255, 255);
    public static Color Gray => new Color(128, 128, 128);
    public static Color Cyan => new Color(0, 255, 255);
    public static Color SlateGray => new Color(112, 128, 144);
    public static Color SandyBrown => new Color(244, 164, 96);
    public static Color Black => new Color(0, 0, 0);
}
