// ErrorLoggerStub.cs -- stands in for NEA.Rendering.ErrorLogger (Rendering/ErrorLogger.cs:9-42 writes to a log file through the
// MonoGame content path); the physics sources only call LogError (RigidBody.cs:93, SATCollision.cs:20).
using System;

namespace NEA.Rendering;

public static class ErrorLogger
{
    public static int Count;

    public static void LogError(string error)
    {
        Count++;
        Console.Error.WriteLine("[ErrorLogger] " + error);
    }
}
