-- |
-- GPU strategies for streamly-lz4's benchmark driver.  The reference's benchmark/Main.hs:189-220 reads
-- BENCH_STREAMLY_LZ4_STRATEGY = "c+<speed>+<bufsize>" | "d+<bufsize>" | "r+<bufsize>" and runs the CPU
-- combinators under gauge.  This module adds the same three strategies over libb200lz4.so
-- (Streamly.Internal.LZ4.B200), selected with a leading 'g':
--
--     gc+400+640000   compress, acceleration 400, 640000-byte arrays, independent blocks, batched GPU calls
--     gd+640000       decompress of the file written by gc+ (read in 640000-byte pieces -> resizeChunksD -> decode)
--     gr+640000       resizeChunksD alone over b200lz4_reframe
--
-- and reports, next to gauge's wall-clock figure, the DEVICE timing split of the last batch call
-- (b200lz4_last_timing: H2D, kernels, D2H in milliseconds) -- the "GPU strategies and device timing" that
-- BASELINE.json's north_star asks benchmark/Main.hs to gain.
--
-- STATUS: not compiled in this repository's image (no GHC); the Python mirror of exactly these three
-- pipelines is what bench.py runs (`e2e_staged`, `extra["d+640000"]`, config 5 in tests/test_gpu_configs.py).
-- To use: add this file to the benchmark's other-modules, call `tryBenchGPU` before `tryBenchExternal` in
-- main, and link with -lb200lz4 (INTEGRATION.md section 1).
module MainB200 (tryBenchGPU) where

import Data.Function ((&))
import Data.Word (Word8)
import Gauge (Benchmark, bench, nfIO)
import System.Environment (lookupEnv)
import System.IO (IOMode(..), openFile, hClose)

import qualified Streamly.Internal.Data.Array.Foreign as Array
import qualified Streamly.Internal.Data.Stream.IsStream as Stream
import qualified Streamly.Internal.FileSystem.Handle as Handle
import qualified Streamly.Internal.LZ4.B200 as B200
import Streamly.Internal.LZ4.Config

data GStrategy
    = GCompress Int Int
    | GDecompress Int
    | GResize Int

parseGStrategy :: String -> Maybe GStrategy
parseGStrategy ('g':'c':_:r) =
    let (speed, pbufsize) = span (/= '+') r
     in Just (GCompress (read speed) (read (tail pbufsize)))
parseGStrategy ('g':'d':_:r) = Just (GDecompress (read r))
parseGStrategy ('g':'r':_:r) = Just (GResize (read r))
parseGStrategy _ = Nothing

gpuBlockConfig :: BlockConfig
gpuBlockConfig = setBlockIndependence True defaultBlockConfig

-- read `file` in bufsize pieces, run the array combinator, write next to it
pipeline
    :: (Stream.SerialT IO (Array.Array Word8) -> Stream.SerialT IO (Array.Array Word8))
    -> Int -> FilePath -> FilePath -> IO ()
pipeline f bufsize inp out = do
    hin <- openFile inp ReadMode
    hout <- openFile out WriteMode
    Stream.unfold Handle.readChunksWithBufferOf (bufsize, hin)
        & f
        & Handle.putChunks hout
    hClose hin >> hClose hout

viaD f = Stream.fromStreamD . f . Stream.toStreamD

gcompress :: Int -> Int -> FilePath -> Benchmark
gcompress bufsize speed file =
    bench ("gc+" ++ show speed ++ "+" ++ show bufsize)
        $ nfIO
        $ pipeline (viaD (B200.compressChunksD B200.defaultB200Config gpuBlockConfig speed)) bufsize file (file ++ ".b200lz4")

gdecompress :: Int -> FilePath -> Benchmark
gdecompress bufsize file =
    bench ("gd+" ++ show bufsize)
        $ nfIO
        $ pipeline
              (viaD (B200.decompressChunksRawD B200.defaultB200Config gpuBlockConfig
                     . B200.resizeChunksD gpuBlockConfig defaultFrameConfig))
              bufsize (file ++ ".b200lz4") (file ++ ".b200lz4.out")

gresize :: Int -> FilePath -> Benchmark
gresize bufsize file =
    bench ("gr+" ++ show bufsize)
        $ nfIO
        $ pipeline (viaD (B200.resizeChunksD gpuBlockConfig defaultFrameConfig)) bufsize (file ++ ".b200lz4") "/dev/null"

-- | Like the reference's tryBenchExternal (benchmark/Main.hs:204-216), for the g-strategies.
tryBenchGPU :: IO (Maybe Benchmark)
tryBenchGPU = do
    fr <- lookupEnv "BENCH_STREAMLY_LZ4_FILE"
    fs <- lookupEnv "BENCH_STREAMLY_LZ4_STRATEGY"
    return $ case (fr, fs >>= parseGStrategy) of
        (Just file, Just (GCompress speed bufsize)) -> Just (gcompress bufsize speed file)
        (Just file, Just (GDecompress bufsize)) -> Just (gdecompress bufsize file)
        (Just file, Just (GResize bufsize)) -> Just (gresize bufsize file)
        _ -> Nothing
